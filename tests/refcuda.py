"""TEST INFRASTRUCTURE: ctypes binding of oracle/_ref/libRefCudaRasterizer.so -- the UNMODIFIED reference
CUDA rasterizer compiled for sm_100a by oracle/Makefile plus the flat shim oracle/ref_shim.cu.
Only tests/, __graft_entry__.smoke() and bench.py's reference legs may import this."""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libRefCudaRasterizer.so")

_p, _i, _f, _z = C.c_void_p, C.c_int, C.c_float, C.c_size_t
_VIEW = [_p, _p, _p, _p, _p, _f, _p, _p, _p, _p, _p, _f, _f]
_lib = None


def available() -> bool:
    return os.path.exists(REF_LIB)


def lib():
    global _lib
    if _lib is None:
        l = C.CDLL(REF_LIB)
        l.ref_required_geom.restype = _z; l.ref_required_geom.argtypes = [_z]
        l.ref_required_image.restype = _z; l.ref_required_image.argtypes = [_z]
        l.ref_required_binning.restype = _z; l.ref_required_binning.argtypes = [_z]
        l.ref_mark_visible.restype = None; l.ref_mark_visible.argtypes = [_i, _p, _p, _p, _p]
        l.ref_forward_0.restype = _i
        l.ref_forward_0.argtypes = [_p, _i, _i, _i, _p, _i, _i, *_VIEW, _i, _p, _i]
        l.ref_forward_1.restype = _i
        l.ref_forward_1.argtypes = [_p, _p, _p, _i, _i, _i, _i, _p, _i, _i, *_VIEW, _i, _p, _p, _i]
        l.ref_forward.restype = _i
        l.ref_forward.argtypes = [_p, _z, _p, _z, _p, _z, _i, _i, _i, _p, _i, _i, *_VIEW, _i, _p, _p, _i, _p]
        l.ref_backward.restype = _i
        l.ref_backward.argtypes = [_i, _i, _i, _i, _p, _i, _i, _p, _p, _p, _p, _f, _p, _p, _p, _p, _p, _f, _f, _p,
                                   _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i]
        l.ref_geom_view.restype = None; l.ref_geom_view.argtypes = [_p, _z, _p]
        l.ref_image_view.restype = None; l.ref_image_view.argtypes = [_p, _z, _p]
        l.ref_binning_view.restype = None; l.ref_binning_view.argtypes = [_p, _z, _p]
        _lib = l
    return _lib


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None or t.numel() == 0 else t.data_ptr()


class RefFrame:
    """One forward of the reference (two-phase, as its Jittor glue drives it) and, optionally, its backward."""

    def __init__(self, bg, means3D, opacities, viewmatrix, projmatrix, campos, tanfovx, tanfovy, H, W, degree,
                 shs=None, colors=None, scales=None, rotations=None, cov3D=None, scale_modifier=1.0, M=None, sync=True):
        l = lib()
        dev = means3D.device
        self.P = P = means3D.shape[0]
        self.H, self.W, self.D = H, W, degree
        self.M = M if M is not None else (shs.shape[1] if shs is not None else 0)
        self.inputs = dict(bg=bg, means3D=means3D, opacities=opacities, viewmatrix=viewmatrix, projmatrix=projmatrix,
                           campos=campos, shs=shs, colors=colors, scales=scales, rotations=rotations, cov3D=cov3D)
        self.tan = (float(tanfovx), float(tanfovy))
        self.scale_modifier = float(scale_modifier)
        self.view_args = (_ptr(means3D), _ptr(shs), _ptr(colors), _ptr(opacities), _ptr(scales), self.scale_modifier,
                          _ptr(rotations), _ptr(cov3D), _ptr(viewmatrix), _ptr(projmatrix), _ptr(campos), *self.tan)
        self.geom = torch.zeros(l.ref_required_geom(P), dtype=torch.uint8, device=dev)
        self.image = torch.zeros(l.ref_required_image(H * W), dtype=torch.uint8, device=dev)
        self.radii = torch.zeros(P, dtype=torch.int32, device=dev)
        self.color = torch.zeros(3, H, W, dtype=torch.float32, device=dev)
        if sync:
            torch.cuda.synchronize()
        R = l.ref_forward_0(self.geom.data_ptr(), P, degree, self.M, _ptr(bg), W, H, *self.view_args, 0,
                            self.radii.data_ptr(), 0)
        assert R >= 0, "reference forward_0 failed"
        self.R = R
        self.binning = torch.zeros(l.ref_required_binning(R), dtype=torch.uint8, device=dev)
        rc = l.ref_forward_1(self.geom.data_ptr(), self.binning.data_ptr(), self.image.data_ptr(), P, degree, self.M, R,
                             _ptr(bg), W, H, *self.view_args, 0, self.color.data_ptr(), self.radii.data_ptr(), 0)
        assert rc == 0, "reference forward_1 failed"
        if sync:
            torch.cuda.synchronize()

    def geom_state(self) -> Dict[str, torch.Tensor]:
        """Copies of the reference's per-Gaussian intermediate state."""
        P = self.P
        ptrs = (C.c_void_p * 9)()
        lib().ref_geom_view(self.geom.data_ptr(), P, ptrs)
        base = self.geom.data_ptr()

        def view(k, dtype, shape):
            off = ptrs[k] - base
            n = int(torch.tensor([], dtype=dtype).element_size())
            cnt = 1
            for s in shape:
                cnt *= s
            return self.geom[off:off + cnt * n].view(dtype).view(*shape).clone()

        return {"depths": view(0, torch.float32, (P,)), "clamped": view(1, torch.bool, (P, 3)),
                "means2D": view(3, torch.float32, (P, 2)), "cov3D": view(4, torch.float32, (P, 6)),
                "conic_opacity": view(5, torch.float32, (P, 4)), "rgb": view(6, torch.float32, (P, 3)),
                "tiles_touched": view(7, torch.int32, (P,))}

    def image_state(self) -> Dict[str, torch.Tensor]:
        N = self.H * self.W
        ptrs = (C.c_void_p * 3)()
        lib().ref_image_view(self.image.data_ptr(), N, ptrs)
        base = self.image.data_ptr()
        o0, o1 = ptrs[0] - base, ptrs[1] - base
        return {"final_T": self.image[o0:o0 + 4 * N].view(torch.float32).view(self.H, self.W).clone(),
                "n_contrib": self.image[o1:o1 + 4 * N].view(torch.int32).view(self.H, self.W).clone()}

    def backward(self, dL_dpix: torch.Tensor, sync=True, M: Optional[int] = None) -> Dict[str, torch.Tensor]:
        """`M` overrides the SH row length of the forward: the reference's backward glue takes it from the SH tensor alone
        (rasterize_points.py:301) even when its forward ran with M = 0 (:154-158)."""
        P, M = self.P, (self.M if M is None else M)
        dev = dL_dpix.device
        i = self.inputs
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
        g = {"means2D": z(P, 3), "conic": z(P, 4), "opacity": z(P, 1), "colors": z(P, 3), "means3D": z(P, 3),
             "cov3D": z(P, 6), "sh": z(P, max(M, 1), 3), "scales": z(P, 3), "rotations": z(P, 4)}
        dL = dL_dpix.contiguous()
        rc = lib().ref_backward(P, self.D, M, self.R, _ptr(i["bg"]), self.W, self.H, _ptr(i["means3D"]), _ptr(i["shs"]),
                                _ptr(i["colors"]), _ptr(i["scales"]), self.scale_modifier, _ptr(i["rotations"]),
                                _ptr(i["cov3D"]), _ptr(i["viewmatrix"]), _ptr(i["projmatrix"]), _ptr(i["campos"]),
                                *self.tan, self.radii.data_ptr(), self.geom.data_ptr(), self.binning.data_ptr(),
                                self.image.data_ptr(), dL.data_ptr(), g["means2D"].data_ptr(), g["conic"].data_ptr(),
                                g["opacity"].data_ptr(), g["colors"].data_ptr(), g["means3D"].data_ptr(),
                                g["cov3D"].data_ptr(), g["sh"].data_ptr(), g["scales"].data_ptr(),
                                g["rotations"].data_ptr(), 0)
        assert rc == 0, "reference backward failed"
        if sync:
            torch.cuda.synchronize()
        if M == 0:
            g["sh"] = g["sh"][:, :0]
        return g


class RefArena:
    """The reference's best case (BASELINE.md 2.1 variant i): its single-call Rasterizer::forward
    (rasterizer_impl.cu:198-336) over PERSISTENT, pre-sized chunks -- no per-frame allocation, no memset of the chunks --
    and its backward into one pre-allocated gradient slab (zeroed once per frame: the reference's backward accumulates
    with atomics).  The mid-frame host read of num_rendered is inside Rasterizer::forward and stays."""

    def __init__(self, device, P: int, H: int, W: int, M: int, binning_instances: int):
        l = lib()
        self.P, self.H, self.W, self.M = P, H, W, M
        self.geom = torch.empty(l.ref_required_geom(P), dtype=torch.uint8, device=device)
        self.image = torch.empty(l.ref_required_image(H * W), dtype=torch.uint8, device=device)
        self.binning = torch.empty(l.ref_required_binning(int(binning_instances)), dtype=torch.uint8, device=device)
        self.radii = torch.empty(P, dtype=torch.int32, device=device)
        self.color = torch.empty(3, H, W, dtype=torch.float32, device=device)
        self._needed = C.c_size_t(0)
        sizes = {"means2D": 3 * P, "conic": 4 * P, "opacity": P, "colors": 3 * P, "means3D": 3 * P, "cov3D": 6 * P,
                 "sh": 3 * max(M, 1) * P, "scales": 3 * P, "rotations": 4 * P}
        total = sum(sizes.values())
        self.slab = torch.empty(total, dtype=torch.float32, device=device)
        self.grads, off = {}, 0
        for k, n in sizes.items():
            self.grads[k] = self.slab[off:off + n]
            off += n
        self.R = 0

    def forward(self, bg, means3D, opacities, viewmatrix, projmatrix, campos, tanfovx, tanfovy, degree, shs=None, colors=None,
                scales=None, rotations=None, cov3D=None, scale_modifier=1.0):
        self._args = (means3D, shs, colors, opacities, scales, float(scale_modifier), rotations, cov3D, viewmatrix, projmatrix,
                      campos, float(tanfovx), float(tanfovy), bg, degree)
        va = (_ptr(means3D), _ptr(shs), _ptr(colors), _ptr(opacities), _ptr(scales), float(scale_modifier), _ptr(rotations),
              _ptr(cov3D), _ptr(viewmatrix), _ptr(projmatrix), _ptr(campos), float(tanfovx), float(tanfovy))
        R = lib().ref_forward(self.geom.data_ptr(), self.geom.numel(), self.binning.data_ptr(), self.binning.numel(),
                              self.image.data_ptr(), self.image.numel(), self.P, degree, self.M, _ptr(bg), self.W, self.H, *va, 0,
                              self.color.data_ptr(), self.radii.data_ptr(), 0, C.byref(self._needed))
        if R == -2:
            raise RuntimeError(f"reference arena too small: binning needs {self._needed.value} bytes")
        assert R >= 0, "reference forward failed"
        self.R = R
        return self.color

    def backward(self, dL_dpix: torch.Tensor) -> Dict[str, torch.Tensor]:
        means3D, shs, colors, opacities, scales, mod, rotations, cov3D, viewmatrix, projmatrix, campos, tx, ty, bg, degree = self._args
        self.slab.zero_()
        g = self.grads
        rc = lib().ref_backward(self.P, degree, self.M, self.R, _ptr(bg), self.W, self.H, _ptr(means3D), _ptr(shs), _ptr(colors),
                                _ptr(scales), mod, _ptr(rotations), _ptr(cov3D), _ptr(viewmatrix), _ptr(projmatrix), _ptr(campos),
                                tx, ty, self.radii.data_ptr(), self.geom.data_ptr(), self.binning.data_ptr(),
                                self.image.data_ptr(), dL_dpix.data_ptr(), g["means2D"].data_ptr(), g["conic"].data_ptr(),
                                g["opacity"].data_ptr(), g["colors"].data_ptr(), g["means3D"].data_ptr(), g["cov3D"].data_ptr(),
                                g["sh"].data_ptr(), g["scales"].data_ptr(), g["rotations"].data_ptr(), 0)
        assert rc == 0, "reference backward failed"
        return g


def deform_gaussians_torch(vertex, vertex_deformed, vertex_R, vertex_S, triangles, weights, pos, cov6):
    """SingleObjectDeform.deform_gaussian (edittool/__init__.py:103-131) as the tensor-op chain the reference runs (torch
    standing in for Jittor).  cov6 packed in / out; returns (pos', cov6', R_g)."""
    tri = triangles.long()
    w_pos = weights[:, :, None]
    w_rs = weights[:, :, None, None]
    g_delta_pos = (w_pos * (vertex_deformed - vertex)[tri]).sum(dim=1)
    g_delta_r = (w_rs * vertex_R[tri]).sum(dim=1)
    rot = g_delta_r.transpose(1, 2)
    g_delta_s = (w_rs * vertex_S[tri]).sum(dim=1)
    g_delta_rs = torch.matmul(rot, g_delta_s)
    c = cov6
    full = torch.stack([c[:, 0], c[:, 1], c[:, 2], c[:, 1], c[:, 3], c[:, 4], c[:, 2], c[:, 4], c[:, 5]], dim=1).view(-1, 3, 3)
    d = torch.matmul(torch.matmul(g_delta_rs, full), g_delta_rs.transpose(1, 2))
    cov_out = torch.stack([d[:, 0, 0], d[:, 0, 1], d[:, 0, 2], d[:, 1, 1], d[:, 1, 2], d[:, 2, 2]], dim=1).contiguous()
    return (pos + g_delta_pos).contiguous(), cov_out, rot.contiguous()


def edit_colors_torch(pos: torch.Tensor, campos: torch.Tensor, rot: torch.Tensor, shs: torch.Tensor, deg: int = 3) -> torch.Tensor:
    """The reference's per-frame edit colours (edittool/__init__.py:442-448 + edittool/sh_utils.py:34-89) as the
    same chain of elementwise / bmm tensor ops the Jittor code runs (torch standing in for Jittor)."""
    C0 = 0.28209479177387814; C1 = 0.4886025119029199
    C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
    C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
          1.445305721320277, -0.5900435899266435]
    sh = shs.transpose(1, 2)
    d = pos - campos.repeat(pos.shape[0], 1)
    d = d / d.norm(dim=1, keepdim=True)
    d = torch.matmul(rot.transpose(1, 2), d.unsqueeze(2)).squeeze(2)
    result = C0 * sh[..., 0]
    if deg > 0:
        x, y, z = d[..., 0:1], d[..., 1:2], d[..., 2:3]
        result = result - C1 * y * sh[..., 1] + C1 * z * sh[..., 2] - C1 * x * sh[..., 3]
        if deg > 1:
            xx, yy, zz = x * x, y * y, z * z
            xy, yz, xz = x * y, y * z, x * z
            result = (result + C2[0] * xy * sh[..., 4] + C2[1] * yz * sh[..., 5] + C2[2] * (2.0 * zz - xx - yy) * sh[..., 6]
                      + C2[3] * xz * sh[..., 7] + C2[4] * (xx - yy) * sh[..., 8])
            if deg > 2:
                result = (result + C3[0] * y * (3 * xx - yy) * sh[..., 9] + C3[1] * xy * z * sh[..., 10]
                          + C3[2] * y * (4 * zz - xx - yy) * sh[..., 11] + C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[..., 12]
                          + C3[4] * x * (4 * zz - xx - yy) * sh[..., 13] + C3[5] * z * (xx - yy) * sh[..., 14]
                          + C3[6] * x * (xx - 3 * yy) * sh[..., 15])
    return torch.clamp(result + 0.5, min=0.0)
