"""TEST INFRASTRUCTURE ONLY -- numpy fp32 restatement of the reference's PYTHON (Jittor) side of the hot path:
the op chains that run before the rasterizer in training (mesh bind + activations, Python covariance / SH
fallbacks) and in editing (per-face deform, rotated-direction colours).  This is "the reference's Jittor CPU
preprocess path" of BASELINE.json: Jittor cannot be installed here, so each function below restates the
Jittor expression it cites with the same op order in numpy float32.

Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may import this module.

Pinning: the reference ships no fixtures for these functions; they are one-line array
expressions transcribed from the cited lines, cross-checked in tests/test_oracle.py against independent
formulations (float64 closed forms, the C rasterizer oracle's SH / covariance code).
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np

f32 = np.float32

C0 = 0.28209479177387814
C1 = 0.4886025119029199
C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
      1.445305721320277, -0.5900435899266435]


def sigmoid(x):
    return (1.0 / (1.0 + np.exp(-x.astype(f32)))).astype(f32)


def softmax(x, axis=1):
    e = np.exp(x - x.max(axis=axis, keepdims=True))
    return (e / e.sum(axis=axis, keepdims=True)).astype(f32)


def get_xyz(bc_logits, distance, vertex1, vertex2, vertex3, normal, r, alpha_distance=4.0):
    """scene/mesh_based_gaussian_model.py:139-152"""
    bc = softmax(bc_logits.astype(f32), axis=1)
    proj_xyz = bc[:, 0:1] * vertex1 + bc[:, 1:2] * vertex2 + bc[:, 2:3] * vertex3
    offset = f32(alpha_distance) * r * (sigmoid(distance) - f32(0.5)) * normal
    return (proj_xyz + offset).astype(f32)


def get_scaling(log_scales):
    """scene/mesh_based_gaussian_model.py:35,122-124"""
    return np.exp(log_scales.astype(f32)).astype(f32)


def get_rotation(rot_raw):
    """scene/mesh_based_gaussian_model.py:43,126-128: jt.normalize = x / max(||x||_2, eps)"""
    n = np.sqrt((rot_raw.astype(f32) ** 2).sum(axis=1, keepdims=True))
    return (rot_raw / np.maximum(n, f32(1e-12))).astype(f32)


def get_opacity(opacity_logit):
    """scene/mesh_based_gaussian_model.py:40,172-174"""
    return sigmoid(opacity_logit)


def build_rotation(r):
    """utils/general_utils.py:74-96 (normalises the quaternion, unlike the CUDA path)"""
    r = r.astype(f32)
    norm = np.sqrt(r[:, 0] * r[:, 0] + r[:, 1] * r[:, 1] + r[:, 2] * r[:, 2] + r[:, 3] * r[:, 3])
    q = r / norm[:, None]
    R = np.zeros((q.shape[0], 3, 3), dtype=f32)
    r_, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R[:, 0, 0] = 1 - 2 * (y * y + z * z)
    R[:, 0, 1] = 2 * (x * y - r_ * z)
    R[:, 0, 2] = 2 * (x * z + r_ * y)
    R[:, 1, 0] = 2 * (x * y + r_ * z)
    R[:, 1, 1] = 1 - 2 * (x * x + z * z)
    R[:, 1, 2] = 2 * (y * z - r_ * x)
    R[:, 2, 0] = 2 * (x * z - r_ * y)
    R[:, 2, 1] = 2 * (y * z + r_ * x)
    R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def strip_symmetric(sym):
    """utils/general_utils.py:64-72 / edittool/general_utils.py:26-37"""
    return np.stack([sym[:, 0, 0], sym[:, 0, 1], sym[:, 0, 2], sym[:, 1, 1], sym[:, 1, 2], sym[:, 2, 2]], axis=1).astype(f32)


def build_covariance_from_scaling_rotation(scaling, scaling_modifier, rotation):
    """scene/mesh_based_gaussian_model.py:28-32 with utils/general_utils.py:98-109"""
    R = build_rotation(rotation)
    L = R * (f32(scaling_modifier) * scaling.astype(f32))[:, None, :]      # R @ diag(s)
    return strip_symmetric(L @ np.swapaxes(L, 1, 2))


def eval_sh(deg, sh, dirs):
    """utils/sh_utils.py:57-112 (degrees 0-3).  sh [..., C, K], dirs [..., 3] -> [..., C]"""
    sh = sh.astype(f32)
    dirs = dirs.astype(f32)
    result = f32(C0) * sh[..., 0]
    if deg > 0:
        x, y, z = dirs[..., 0:1], dirs[..., 1:2], dirs[..., 2:3]
        result = result - f32(C1) * y * sh[..., 1] + f32(C1) * z * sh[..., 2] - f32(C1) * x * sh[..., 3]
        if deg > 1:
            xx, yy, zz = x * x, y * y, z * z
            xy, yz, xz = x * y, y * z, x * z
            result = (result + f32(C2[0]) * xy * sh[..., 4] + f32(C2[1]) * yz * sh[..., 5]
                      + f32(C2[2]) * (f32(2.0) * zz - xx - yy) * sh[..., 6] + f32(C2[3]) * xz * sh[..., 7]
                      + f32(C2[4]) * (xx - yy) * sh[..., 8])
            if deg > 2:
                result = (result + f32(C3[0]) * y * (3 * xx - yy) * sh[..., 9] + f32(C3[1]) * xy * z * sh[..., 10]
                          + f32(C3[2]) * y * (4 * zz - xx - yy) * sh[..., 11]
                          + f32(C3[3]) * z * (2 * zz - 3 * xx - 3 * yy) * sh[..., 12]
                          + f32(C3[4]) * x * (4 * zz - xx - yy) * sh[..., 13] + f32(C3[5]) * z * (xx - yy) * sh[..., 14]
                          + f32(C3[6]) * x * (xx - 3 * yy) * sh[..., 15])
    return result.astype(f32)


def sh_to_rgb(deg, shs, xyz, campos, rot=None):
    """convert_SHs_python branch, gaussian_renderer/__init__.py:87-92; with `rot` the edit-time variant that
    rotates the direction into the rest frame first, edittool/__init__.py:442-448.  shs [P,16,3]."""
    shs_view = np.swapaxes(shs.astype(f32), 1, 2)                    # [P,3,16]
    dir_pp = xyz.astype(f32) - campos.astype(f32)[None, :]
    dir_pp_normalized = dir_pp / np.sqrt((dir_pp ** 2).sum(axis=1, keepdims=True))
    if rot is not None:
        dir_pp_normalized = (np.swapaxes(rot.astype(f32), 1, 2) @ dir_pp_normalized[:, :, None])[:, :, 0]
    sh2rgb = eval_sh(deg, shs_view, dir_pp_normalized)
    return np.maximum(sh2rgb + f32(0.5), f32(0.0)).astype(f32)


def get_barycentric_coordinate(gaussians, p1, p2, p3):
    """edittool/general_utils.py:73-88 (numpy float64 in the reference)"""
    e1, e2, e3 = gaussians - p1, gaussians - p2, gaussians - p3
    s1 = np.linalg.norm(np.cross(e2, e3), axis=1)[:, None]
    s2 = np.linalg.norm(np.cross(e1, e3), axis=1)[:, None]
    s3 = np.linalg.norm(np.cross(e1, e2), axis=1)[:, None]
    s = s1 + s2 + s3
    return np.concatenate([s1 / s, s2 / s, s3 / s], axis=1)


def deform_gaussian(vertex, deform_vertex, R1, S1, gaussian_triangles, coord, gaussian_pos, gaussian_cov
                    ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """SingleObjectDeform.deform_gaussian, edittool/__init__.py:103-131.
    vertex/deform_vertex [Vn,3]; R1/S1 [Vn,3,3] (ACAP rotation / shear); gaussian_triangles [P,3] vertex ids;
    coord [P,3] barycentric weights; gaussian_cov [P,3,3].  Returns (pos', cov' [P,3,3], R_g [P,3,3])."""
    weight_g_pos = coord.astype(f32)[:, :, None]
    weight_g_rs = coord.astype(f32)[:, :, None, None]
    cur_pos, cur_rot, cur_shear = deform_vertex.astype(f32), R1.astype(f32), S1.astype(f32)
    delta_pos_ = (cur_pos - vertex.astype(f32))[gaussian_triangles]
    g_delta_pos = np.sum(weight_g_pos * delta_pos_, axis=1)
    R_ = cur_rot[gaussian_triangles]
    g_delta_r = np.sum(weight_g_rs * R_, axis=1)
    gaussian_deform_rot = np.swapaxes(g_delta_r, 1, 2)
    S_ = cur_shear[gaussian_triangles]
    g_delta_s = np.sum(weight_g_rs * S_, axis=1)
    g_delta_rs = gaussian_deform_rot @ g_delta_s
    gaussian_deform_cov = (g_delta_rs @ gaussian_cov.astype(f32)) @ np.swapaxes(g_delta_rs, 1, 2)
    gaussian_deform_pos = gaussian_pos.astype(f32) + g_delta_pos
    return gaussian_deform_pos.astype(f32), gaussian_deform_cov.astype(f32), gaussian_deform_rot.astype(f32)


def l1_loss(network_output, gt):
    """utils/loss_utils.py:17-18"""
    return np.abs(network_output.astype(f32) - gt.astype(f32)).mean(dtype=np.float64)


def mesh_bound_inputs(arrays: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """The rasterizer inputs of a mesh-bound scene as the reference's accessors compute them."""
    return {
        "means3D": get_xyz(arrays["bc_logits"], arrays["distance"], arrays["vertex1"], arrays["vertex2"], arrays["vertex3"],
                           arrays["normal"], arrays["r"]),
        "scales": get_scaling(arrays["log_scales"]),
        "rotations": get_rotation(arrays["rot_raw"]),
        "opacities": get_opacity(arrays["opacity_logit"]),
        "shs": arrays["shs"],
    }
