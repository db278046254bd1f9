"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's training-loop pieces around the rasterizer op
(SURVEY.md 8f-4): L1 / SSIM / mesh-restrict losses, the learning-rate schedule, Adam and the densification
statistics.  Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may import this module.

The reference writes these as Jittor expressions; Jittor cannot be installed here.  Each function restates the
cited lines op for op -- conv2d-based ones with torch's CPU conv2d (same semantics as jt.nn.conv2d: cross-correlation,
zero padding, groups), gradients by torch autograd (what Jittor's tape computes), the rest in numpy float32.

PARITY UNPINNED by the reference: it ships no tests or fixtures for any of these.  Adam lives in
Jittor itself, which /root/reference neither vendors nor pins (README asks for "jittor"); `adam_step` restates the
update its optim.Adam.step publishes (m, v moments; step size lr*sqrt(1-b1^n)/(1-b0^n); eps added to sqrt(v)).
tests/test_oracle.py cross-checks every function against an independent formulation (scipy separable filtering in
float64 for SSIM, torch.optim.Adam for the update, closed forms for the rest).
"""
from __future__ import annotations

from math import exp
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

f32 = np.float32


# ---- utils/loss_utils.py ----------------------------------------------------------------------
def l1_loss(network_output: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
    """utils/loss_utils.py:17-18"""
    return torch.abs(network_output - gt).mean()


def gaussian(window_size: int, sigma: float) -> torch.Tensor:
    """utils/loss_utils.py:23-25"""
    gauss = torch.tensor([exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)],
                         dtype=torch.float32)
    return gauss / gauss.sum()


def create_window(window_size: int, channel: int) -> torch.Tensor:
    """utils/loss_utils.py:28-33"""
    _1D_window = gaussian(window_size, 1.5).unsqueeze(1)
    _2D_window = _1D_window.mm(_1D_window.t()).float().unsqueeze(0).unsqueeze(0)
    return _2D_window.expand(channel, 1, window_size, window_size).contiguous()


def ssim(img1: torch.Tensor, img2: torch.Tensor, window_size: int = 11, size_average: bool = True) -> torch.Tensor:
    """utils/loss_utils.py:36-82"""
    channel = img1.shape[-3]
    window = create_window(window_size, channel).type_as(img1)
    if img1.dim() == 3:
        img1 = img1.unsqueeze(0)
    if img2.dim() == 3:
        img2 = img2.unsqueeze(0)
    pad = window_size // 2
    mu1 = F.conv2d(img1, window, padding=pad, groups=channel)
    mu2 = F.conv2d(img2, window, padding=pad, groups=channel)
    mu1_sq = mu1.pow(2)
    mu2_sq = mu2.pow(2)
    mu1_mu2 = mu1 * mu2
    sigma1_sq = F.conv2d(img1 * img1, window, padding=pad, groups=channel) - mu1_sq
    sigma2_sq = F.conv2d(img2 * img2, window, padding=pad, groups=channel) - mu2_sq
    sigma12 = F.conv2d(img1 * img2, window, padding=pad, groups=channel) - mu1_mu2
    C1 = 0.01 ** 2
    C2 = 0.03 ** 2
    ssim_map = ((2 * mu1_mu2 + C1) * (2 * sigma12 + C2)) / ((mu1_sq + mu2_sq + C1) * (sigma1_sq + sigma2_sq + C2))
    if size_average:
        return ssim_map.mean()
    return ssim_map.mean(1).mean(1).mean(1)


def photometric_loss(image: torch.Tensor, gt: torch.Tensor, lambda_dssim: float) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """train_mesh_gaussian.py:91,94 without mrloss: (loss, Ll1, ssim)"""
    Ll1 = l1_loss(image, gt)
    s = ssim(image, gt)
    return (1.0 - lambda_dssim) * Ll1 + lambda_dssim * (1.0 - s), Ll1, s


def photometric_loss_and_grad(image: np.ndarray, gt: np.ndarray, lambda_dssim: float, dtype=torch.float32):
    """(loss, L1, SSIM, dloss/dimage) as numpy; dtype float64 gives the tight reference for tolerance studies."""
    x = torch.from_numpy(np.ascontiguousarray(image)).to(dtype).requires_grad_(True)
    y = torch.from_numpy(np.ascontiguousarray(gt)).to(dtype)
    loss, l1, s = photometric_loss(x, y, lambda_dssim)
    loss.backward()
    return float(loss.detach()), float(l1.detach()), float(s.detach()), x.grad.numpy()


def circumradius(point1: np.ndarray, point2: np.ndarray, point3: np.ndarray) -> np.ndarray:
    """utils/loss_utils.py:87-100 (the "radius" is sqrt of the parallelogram area)"""
    AB = point2 - point1
    AC = point3 - point1
    cross_product = np.cross(AB, AC, axis=1)
    areas = np.linalg.norm(cross_product, axis=1)
    return np.sqrt(areas).astype(f32)


def mesh_restrict_loss(scale: np.ndarray, point1, point2, point3, weight: float = 10) -> Tuple[float, np.ndarray]:
    """utils/loss_utils.py:102-107; returns (loss, dloss/dscale)"""
    max_s = scale.max(axis=1)
    r = circumradius(point1.astype(f32), point2.astype(f32), point3.astype(f32))
    loss = max_s - f32(weight) * r
    on = loss > 0
    grad = np.zeros_like(scale, dtype=f32)
    grad[np.arange(scale.shape[0]), scale.argmax(axis=1)] = on.astype(f32)
    return float(np.clip(loss, 0, None).astype(np.float64).sum()), grad


# ---- utils/general_utils.py:29-62 --------------------------------------------------------------
def get_expon_lr_func(lr_init, lr_final, lr_delay_steps=0, lr_delay_mult=1.0, max_steps=1000000):
    def helper(step):
        if step < 0 or (lr_init == 0.0 and lr_final == 0.0):
            return 0.0
        if lr_delay_steps > 0:
            delay_rate = lr_delay_mult + (1 - lr_delay_mult) * np.sin(0.5 * np.pi * np.clip(step / lr_delay_steps, 0, 1))
        else:
            delay_rate = 1.0
        t = np.clip(step / max_steps, 0, 1)
        log_lerp = np.exp(np.log(lr_init) * (1 - t) + np.log(lr_final) * t)
        return delay_rate * log_lerp
    return helper


# ---- jittor.optim.Adam.step as used by scene/mesh_based_gaussian_model.py:258 ---------------------
def adam_step(p: np.ndarray, g: np.ndarray, m: np.ndarray, v: np.ndarray, lr, n: int, b0: float = 0.9,
              b1: float = 0.999, eps: float = 1e-15) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """One update; `lr` is a scalar or an array broadcastable to p (per-element rates).  float32 like the device."""
    p, g, m, v = (a.astype(f32) for a in (p, g, m, v))
    m = f32(b0) * m + f32(1.0 - b0) * g
    v = f32(b1) * v + f32(1.0 - b1) * g * g
    step_size = (np.asarray(lr, dtype=np.float64) * np.sqrt(1.0 - b1 ** n) / (1.0 - b0 ** n)).astype(f32)
    p = p - m * step_size / (np.sqrt(v) + f32(eps))
    return p.astype(f32), m.astype(f32), v.astype(f32)


# ---- train_mesh_gaussian.py:117-121, scene/mesh_based_gaussian_model.py:587-589 ------------------
def densify_stats(radii: np.ndarray, viewspace_grad: np.ndarray, max_radii2D: np.ndarray, grad_accum: np.ndarray,
                  denom: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    vis = radii > 0
    max_radii2D = max_radii2D.copy()
    grad_accum = grad_accum.copy()
    denom = denom.copy()
    max_radii2D[vis] = np.maximum(max_radii2D[vis], radii[vis].astype(f32))
    grad_accum[vis] += np.linalg.norm(viewspace_grad[vis, :2], axis=-1, keepdims=True).reshape(grad_accum[vis].shape)
    denom[vis] += 1
    return max_radii2D, grad_accum, denom


# ---- scene/mesh_based_gaussian_model.py:504-585, utils/general_utils.py:133-212 ----------------------
def split_mesh_and_gaussian(new_vertex1, new_vertex2, new_vertex3, new_v, new_v_index, v_origin_num, N=4):
    """utils/general_utils.py:133-170 (N = 4) and :172-212 (N = 5); arrays are [S,N,3] ([S,3,3] for new_v)."""
    a = new_vertex1[:, 0, :].copy()
    b = new_vertex2[:, 0, :].copy()
    c = new_vertex3[:, 0, :].copy()
    new_vertex1[:, 0, :] = a
    new_vertex1[:, 1, :] = (a + b) / 2
    new_vertex1[:, 2, :] = (a + c) / 2
    new_vertex1[:, 3, :] = (a + b) / 2
    new_vertex2[:, 0, :] = (a + b) / 2
    new_vertex2[:, 1, :] = b
    new_vertex2[:, 2, :] = (c + b) / 2
    new_vertex2[:, 3, :] = (b + c) / 2
    new_vertex3[:, 0, :] = (a + c) / 2
    new_vertex3[:, 1, :] = (c + b) / 2
    new_vertex3[:, 2, :] = c
    new_vertex3[:, 3, :] = (a + c) / 2
    if N == 5:
        new_vertex1[:, 4, :] = a
        new_vertex2[:, 4, :] = b
        new_vertex3[:, 4, :] = c
    new_v[:, 0, :] = (a + b) / 2
    new_v[:, 1, :] = (a + c) / 2
    new_v[:, 2, :] = (b + c) / 2
    tmp = np.arange(new_v.shape[0] * 3).reshape(new_v.shape[0], 3)
    new_v_index[:, 0, 1] = tmp[:, 0] + v_origin_num
    new_v_index[:, 0, 2] = tmp[:, 1] + v_origin_num
    new_v_index[:, 1, 0] = tmp[:, 0] + v_origin_num
    new_v_index[:, 1, 2] = tmp[:, 2] + v_origin_num
    new_v_index[:, 2, 0] = tmp[:, 1] + v_origin_num
    new_v_index[:, 2, 1] = tmp[:, 2] + v_origin_num
    new_v_index[:, 3, 0] = tmp[:, 0] + v_origin_num
    new_v_index[:, 3, 1] = tmp[:, 2] + v_origin_num
    new_v_index[:, 3, 2] = tmp[:, 1] + v_origin_num
    return new_vertex1, new_vertex2, new_vertex3, new_v, new_v_index


def densify_and_split(model: Dict[str, np.ndarray], moments: Dict[str, Tuple[np.ndarray, np.ndarray]], grads: np.ndarray,
                      grad_threshold: float, N: int = 4):
    """scene/mesh_based_gaussian_model.py:504-563 on a dict of numpy arrays (bc, distance, f_dc, f_rest, opacity,
    scaling, rotation: the parameters; vertex1..3, normal, r, fid, vertex_index: per-Gaussian constants; v: the mesh
    vertices).  `moments[name]` = (m, values) of the seven parameter groups.  Returns (model', moments', selected)."""
    sel = grads.reshape(-1) >= grad_threshold
    if sel.sum() == 0:
        return model, moments, sel
    S = int(sel.sum())
    rep = lambda x: np.repeat(x[sel][:, None], N, axis=1)
    gaussian_num = S * N
    new_bc = np.ones((gaussian_num, 3), f32) / 3
    new_distance = np.zeros((gaussian_num, 1), f32)
    new_v_index = rep(model["vertex_index"])
    new_v = np.zeros((S, 3, 3), f32)
    nv1, nv2, nv3, new_v, new_v_index = split_mesh_and_gaussian(rep(model["vertex1"]), rep(model["vertex2"]),
                                                                 rep(model["vertex3"]), new_v, new_v_index,
                                                                 model["v"].shape[0], N)
    new = {
        "bc": new_bc, "distance": new_distance,
        "f_dc": rep(model["f_dc"]).reshape(gaussian_num, -1, 3), "f_rest": rep(model["f_rest"]).reshape(gaussian_num, -1, 3),
        "opacity": rep(model["opacity"]).reshape(gaussian_num, -1),
        "scaling": np.log(np.exp(model["scaling"].astype(f32))[sel][:, None].repeat(N, axis=1).reshape(gaussian_num, 3) / f32(4 * 0.8)).astype(f32),
        "rotation": rep(model["rotation"]).reshape(gaussian_num, 4),
        "vertex1": nv1.reshape(gaussian_num, 3), "vertex2": nv2.reshape(gaussian_num, 3), "vertex3": nv3.reshape(gaussian_num, 3),
        "vertex_index": new_v_index.reshape(gaussian_num, 3),
        "r": rep(model["r"]).reshape(gaussian_num, -1), "fid": rep(model["fid"]).reshape(gaussian_num, -1),
        "normal": rep(model["normal"]).reshape(gaussian_num, 3),
    }
    keep = ~sel                                   # concat, then prune_points(selected) (:547-563)
    out = {k: np.concatenate([model[k][keep], new[k]], axis=0) for k in new}
    out["v"] = np.concatenate([model["v"], new_v.reshape(-1, 3)], axis=0)
    mom = {k: (np.concatenate([m[keep], np.zeros_like(new[k])], axis=0), np.concatenate([v[keep], np.zeros_like(new[k])], axis=0))
           for k, (m, v) in moments.items()}
    return out, mom, sel


def reset_opacity(opacity_logit: np.ndarray) -> np.ndarray:
    """scene/mesh_based_gaussian_model.py:334-339"""
    x = np.minimum(1.0 / (1.0 + np.exp(-opacity_logit.astype(f32))), f32(0.01)).astype(f32)
    return np.log(x / (1 - x)).astype(f32)
