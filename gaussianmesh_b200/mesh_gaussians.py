"""Host side of the mesh-bound Gaussian kernels (torch + ctypes over libCudaRasterizer.so).

  mesh_bind            training-time per-face bind + activations, differentiable
                       (scene/mesh_based_gaussian_model.py:34-43,122-152)
  deform_gaussians     edit-time per-face local-frame transform (edittool/__init__.py:103-131)
  sh_to_rgb_rotated    edit-time per-frame colour in the rotated frame (edittool/__init__.py:442-448)
  l1_loss              utils/loss_utils.py:17-18, differentiable

No CPU path: CUDA tensors only.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from ._lib import lib, check, RasterizerError, GM_ERR_BAD_ARGUMENT


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _c(t: Optional[torch.Tensor], dtype=torch.float32) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if not t.is_cuda:
        raise RasterizerError("mesh_gaussians", GM_ERR_BAD_ARGUMENT, "tensor is not on a CUDA device (there is no CPU path)")
    if t.dtype != dtype or not t.is_contiguous():
        t = t.to(dtype).contiguous()
    return t


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class _MeshBind(torch.autograd.Function):
    @staticmethod
    def forward(ctx, bc_logits, distance, log_scale, rot_raw, opacity_logit, vertex1, vertex2, vertex3, normal, r,
                alpha_distance):
        bc_logits, distance, log_scale, rot_raw, opacity_logit = map(_c, (bc_logits, distance, log_scale, rot_raw, opacity_logit))
        vertex1, vertex2, vertex3, normal, r = map(_c, (vertex1, vertex2, vertex3, normal, r))
        P = bc_logits.shape[0]
        dev = bc_logits.device
        xyz = torch.empty(P, 3, dtype=torch.float32, device=dev)
        scale = torch.empty(P, 3, dtype=torch.float32, device=dev)
        rot = torch.empty(P, 4, dtype=torch.float32, device=dev)
        opacity = torch.empty(P, 1, dtype=torch.float32, device=dev)
        check(lib.gm_mesh_bind_forward(P, _p(bc_logits), _p(distance), _p(vertex1), _p(vertex2), _p(vertex3), _p(normal),
                                       _p(r), float(alpha_distance), _p(log_scale), _p(rot_raw), _p(opacity_logit),
                                       _p(xyz), _p(scale), _p(rot), _p(opacity), _stream()), "gm_mesh_bind_forward")
        ctx.save_for_backward(bc_logits, distance, log_scale, rot_raw, opacity_logit, vertex1, vertex2, vertex3, normal, r)
        ctx.alpha_distance = float(alpha_distance)
        return xyz, scale, rot, opacity

    @staticmethod
    def backward(ctx, g_xyz, g_scale, g_rot, g_opacity):
        bc_logits, distance, log_scale, rot_raw, opacity_logit, vertex1, vertex2, vertex3, normal, r = ctx.saved_tensors
        P = bc_logits.shape[0]
        dev = bc_logits.device
        g_xyz, g_scale, g_rot, g_opacity = map(_c, (g_xyz, g_scale, g_rot, g_opacity))
        d_bc = torch.zeros(P, 3, dtype=torch.float32, device=dev) if g_xyz is None else torch.empty(P, 3, dtype=torch.float32, device=dev)
        d_dist = torch.zeros(P, 1, dtype=torch.float32, device=dev) if g_xyz is None else torch.empty(P, 1, dtype=torch.float32, device=dev)
        d_ls = torch.zeros(P, 3, dtype=torch.float32, device=dev) if g_scale is None else torch.empty(P, 3, dtype=torch.float32, device=dev)
        d_rr = torch.zeros(P, 4, dtype=torch.float32, device=dev) if g_rot is None else torch.empty(P, 4, dtype=torch.float32, device=dev)
        d_ol = torch.zeros(P, 1, dtype=torch.float32, device=dev) if g_opacity is None else torch.empty(P, 1, dtype=torch.float32, device=dev)
        check(lib.gm_mesh_bind_backward(P, _p(bc_logits), _p(distance), _p(vertex1), _p(vertex2), _p(vertex3), _p(normal),
                                        _p(r), ctx.alpha_distance, _p(log_scale), _p(rot_raw), _p(opacity_logit),
                                        _p(g_xyz), _p(g_scale), _p(g_rot), _p(g_opacity),
                                        _p(d_bc), _p(d_dist), _p(d_ls), _p(d_rr), _p(d_ol), _stream()),
              "gm_mesh_bind_backward")
        return d_bc, d_dist, d_ls, d_rr, d_ol, None, None, None, None, None, None


def mesh_bind(bc_logits, distance, log_scale, rot_raw, opacity_logit, vertex1, vertex2, vertex3, normal, r,
              alpha_distance: float = 4.0) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """(get_xyz, get_scaling, get_rotation, get_opacity) of MeshBasedGaussianModel in one kernel:
        xyz = softmax(bc) . (v1,v2,v3) + alpha_distance * r * (sigmoid(distance) - 0.5) * normal
    Differentiable w.r.t. the five parameter tensors."""
    return _MeshBind.apply(bc_logits, distance, log_scale, rot_raw, opacity_logit, vertex1, vertex2, vertex3, normal,
                           r, alpha_distance)


class _Activate(torch.autograd.Function):
    """exp / normalise / sigmoid of a plain GaussianModel (scene/gaussian_model.py:27-41,96-116) in one kernel."""

    @staticmethod
    def forward(ctx, log_scale, rot_raw, opacity_logit):
        log_scale, rot_raw, opacity_logit = map(_c, (log_scale, rot_raw, opacity_logit))
        P = log_scale.shape[0]
        dev = log_scale.device
        scale = torch.empty(P, 3, dtype=torch.float32, device=dev)
        rot = torch.empty(P, 4, dtype=torch.float32, device=dev)
        opacity = torch.empty(P, 1, dtype=torch.float32, device=dev)
        check(lib.gm_mesh_bind_forward(P, None, None, None, None, None, None, None, 0.0, _p(log_scale), _p(rot_raw),
                                       _p(opacity_logit), None, _p(scale), _p(rot), _p(opacity), _stream()),
              "gm_mesh_bind_forward")
        ctx.save_for_backward(log_scale, rot_raw, opacity_logit)
        return scale, rot, opacity

    @staticmethod
    def backward(ctx, g_scale, g_rot, g_opacity):
        log_scale, rot_raw, opacity_logit = ctx.saved_tensors
        P = log_scale.shape[0]
        dev = log_scale.device
        g_scale, g_rot, g_opacity = map(_c, (g_scale, g_rot, g_opacity))
        alloc = lambda g, n: (torch.zeros if g is None else torch.empty)(P, n, dtype=torch.float32, device=dev)
        d_ls, d_rr, d_ol = alloc(g_scale, 3), alloc(g_rot, 4), alloc(g_opacity, 1)
        check(lib.gm_mesh_bind_backward(P, None, None, None, None, None, None, None, 0.0, _p(log_scale), _p(rot_raw),
                                        _p(opacity_logit), None, _p(g_scale), _p(g_rot), _p(g_opacity), None, None,
                                        _p(d_ls), _p(d_rr), _p(d_ol), _stream()), "gm_mesh_bind_backward")
        return d_ls, d_rr, d_ol


def activate(log_scale, rot_raw, opacity_logit):
    """(get_scaling, get_rotation, get_opacity) of a plain GaussianModel, fused and differentiable."""
    return _Activate.apply(log_scale, rot_raw, opacity_logit)


def deform_gaussians(vertex_rest, vertex_deformed, vertex_R, vertex_S, gaussian_triangles, weights, pos, cov,
                     want_rot: bool = True):
    """SingleObjectDeform.deform_gaussian.  `cov` is [P,6] packed or [P,3,3] full.
    Returns (pos' [P,3], cov' [P,6] packed = strip_symmetric(A Sigma A^T), R_g [P,3,3] or None)."""
    vertex_rest, vertex_deformed, vertex_R, vertex_S, weights, pos, cov = map(
        _c, (vertex_rest, vertex_deformed, vertex_R, vertex_S, weights, pos, cov))
    tri = _c(gaussian_triangles, torch.int32)
    P = pos.shape[0]
    dev = pos.device
    full = 1 if cov.dim() == 3 else 0
    pos_out = torch.empty(P, 3, dtype=torch.float32, device=dev)
    cov_out = torch.empty(P, 6, dtype=torch.float32, device=dev)
    rot_out = torch.empty(P, 3, 3, dtype=torch.float32, device=dev) if want_rot else None
    check(lib.gm_deform_gaussians(P, vertex_rest.shape[0], _p(vertex_rest), _p(vertex_deformed), _p(vertex_R),
                                  _p(vertex_S), _p(tri), _p(weights), _p(pos), _p(cov), full, _p(pos_out),
                                  _p(cov_out), _p(rot_out), _stream()), "gm_deform_gaussians")
    return pos_out, cov_out, rot_out


class _ShToRgb(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pos, shs, campos, rot, degree):
        pos, campos, rot, shs = map(_c, (pos, campos, rot, shs))
        P = pos.shape[0]
        rgb = torch.empty(P, 3, dtype=torch.float32, device=pos.device)
        check(lib.gm_sh_to_rgb_rotated(P, int(degree), int(shs.shape[1]), _p(pos), _p(campos), _p(rot), _p(shs), _p(rgb),
                                       _stream()), "gm_sh_to_rgb_rotated")
        ctx.save_for_backward(pos, shs, campos, rot if rot is not None else torch.empty(0, device=pos.device))
        ctx.degree, ctx.has_rot = int(degree), rot is not None
        return rgb

    @staticmethod
    def backward(ctx, g_rgb):
        pos, shs, campos, rot = ctx.saved_tensors
        rot = rot if ctx.has_rot else None
        g_rgb = _c(g_rgb)
        P = pos.shape[0]
        need_pos, need_shs = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        d_pos = torch.empty_like(pos) if need_pos else None
        d_shs = torch.empty_like(shs) if need_shs else None
        if need_pos or need_shs:
            check(lib.gm_sh_to_rgb_rotated_backward(P, ctx.degree, int(shs.shape[1]), _p(pos), _p(campos), _p(rot), _p(shs),
                                                    _p(g_rgb), _p(d_shs), _p(d_pos), _stream()),
                  "gm_sh_to_rgb_rotated_backward")
        return d_pos, d_shs, None, None, None


def sh_to_rgb_rotated(pos, campos, rot, shs, degree: int = 3) -> torch.Tensor:
    """clamp(eval_sh(degree, shs, R_g^T normalize(pos - campos)) + 0.5, 0); rot may be None (then this is the
    reference's convert_SHs_python branch, gaussian_renderer/__init__.py:87-92).  Differentiable w.r.t. pos and shs."""
    return _ShToRgb.apply(pos, shs, campos, rot, degree)


class _Cov3DPython(torch.autograd.Function):
    @staticmethod
    def forward(ctx, scales, rotations, modifier):
        scales, rotations = _c(scales), _c(rotations)
        P = scales.shape[0]
        cov6 = torch.empty(P, 6, dtype=torch.float32, device=scales.device)
        check(lib.gm_cov3d_from_scale_rot(P, _p(scales), float(modifier), _p(rotations), _p(cov6), _stream()),
              "gm_cov3d_from_scale_rot")
        ctx.save_for_backward(scales, rotations)
        ctx.modifier = float(modifier)
        return cov6

    @staticmethod
    def backward(ctx, g_cov6):
        scales, rotations = ctx.saved_tensors
        g_cov6 = _c(g_cov6)
        d_s, d_r = torch.empty_like(scales), torch.empty_like(rotations)
        check(lib.gm_cov3d_from_scale_rot_backward(scales.shape[0], _p(scales), ctx.modifier, _p(rotations), _p(g_cov6),
                                                   _p(d_s), _p(d_r), _stream()), "gm_cov3d_from_scale_rot_backward")
        return d_s, d_r, None


def covariance_from_scaling_rotation(scaling, scaling_modifier: float, rotation) -> torch.Tensor:
    """pc.get_covariance(scaling_modifier) of the compute_cov3D_python branch: strip_symmetric(L L^T) with
    L = R(rotation / |rotation|) diag(scaling_modifier * scaling) (scene/mesh_based_gaussian_model.py:24-29,
    utils/general_utils.py:64-109) -> [P,6].  Differentiable w.r.t. scaling and the RAW rotation."""
    return _Cov3DPython.apply(scaling, rotation, scaling_modifier)


def load_mesh(vertex, faces, face_id, proj_pos) -> Tuple[torch.Tensor, torch.Tensor]:
    """SingleObjectDeform.load_mesh, face-id branch (edittool/__init__.py:87-101): (gaussian_triangles [P,3] int32,
    barycentric weights [P,3] float64) on the device.  vertex [Vn,3] (float64 like igl.read_triangle_mesh), faces [Fn,3],
    face_id [P] or [P,1], proj_pos [P,3] = get_proj_xyz."""
    proj_pos = _c(proj_pos)
    dev = proj_pos.device
    vertex = torch.as_tensor(vertex, dtype=torch.float64).contiguous().to(dev)
    faces = torch.as_tensor(faces).to(torch.int32).contiguous().to(dev)
    face_id = torch.as_tensor(face_id).to(torch.int64).reshape(-1).contiguous().to(dev)
    P = proj_pos.shape[0]
    if face_id.shape[0] != P:
        raise RasterizerError("load_mesh", GM_ERR_BAD_ARGUMENT, "one face id per Gaussian expected")
    tri = torch.empty(P, 3, dtype=torch.int32, device=dev)
    w = torch.empty(P, 3, dtype=torch.float64, device=dev)
    check(lib.gm_load_mesh(P, int(vertex.shape[0]), int(faces.shape[0]), _p(vertex), _p(faces), _p(face_id), _p(proj_pos),
                           _p(tri), _p(w), _stream()), "gm_load_mesh")
    return tri, w


class _L1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, target):
        img = _c(img)
        u8 = target.dtype == torch.uint8
        target = _c(target, torch.uint8 if u8 else torch.float32)
        loss = torch.empty(1, dtype=torch.float32, device=img.device)
        grad = torch.empty_like(img)
        fn = lib.gm_l1_loss_u8 if u8 else lib.gm_l1_loss
        check(fn(img.numel(), _p(img), _p(target), _p(loss), _p(grad), _stream()), "gm_l1_loss")
        ctx.save_for_backward(grad)
        return loss.view(())

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None


def image_u8_to_float(img_u8: torch.Tensor) -> torch.Tensor:
    """uint8 image -> float32 / 255.0 (utils/general_utils.py:22-27, PILtoTorch), one kernel."""
    img_u8 = _c(img_u8, torch.uint8)
    out = torch.empty(img_u8.shape, dtype=torch.float32, device=img_u8.device)
    check(lib.gm_image_u8_to_float(img_u8.numel(), _p(img_u8), _p(out), _stream()), "gm_image_u8_to_float")
    return out


def l1_loss(network_output: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
    """mean(abs(network_output - gt)) (utils/loss_utils.py:17-18), fused with its gradient.  `gt` may be the uint8 image
    (value / 255 inside the kernel)."""
    return _L1.apply(network_output, gt)
