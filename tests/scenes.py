"""Shared builders for the parity tests: seeded synthetic inputs as torch tensors."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np
import torch

from gaussianmesh_b200 import synthetic


def to_dev(arrays: Dict[str, np.ndarray], device) -> Dict[str, torch.Tensor]:
    return {k: torch.from_numpy(np.ascontiguousarray(v)).to(device) for k, v in arrays.items()}


def free_scene(P: int, device, seed: int = 0, **kw) -> Dict[str, torch.Tensor]:
    return to_dev(synthetic.gaussian_scene(P, seed=seed, **kw), device)


def camera(device, W: int, H: int, index: int = 0, n: int = 7, **kw):
    from gaussianmesh_b200.renderer import DeviceCamera
    cam = synthetic.orbit_cameras(n, W, H, **kw)[index]
    return DeviceCamera.upload(cam, device)


def packed_cov(scales: torch.Tensor, rotations: torch.Tensor, modifier: float = 1.0) -> torch.Tensor:
    """build_covariance_from_scaling_rotation + strip_symmetric (utils/general_utils.py:64-109) in torch,
    fp64 then rounded -- only used to create precomputed-covariance INPUTS."""
    s = scales.double() * modifier
    q = rotations.double()
    q = q / q.norm(dim=1, keepdim=True)
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
        2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
        2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], dim=1).view(-1, 3, 3)
    L = R * s[:, None, :]
    S = L @ L.transpose(1, 2)
    return torch.stack([S[:, 0, 0], S[:, 0, 1], S[:, 0, 2], S[:, 1, 1], S[:, 1, 2], S[:, 2, 2]], dim=1).float().contiguous()


def our_geom_state(geom: torch.Tensor, P: int, num_tiles: int) -> Dict[str, torch.Tensor]:
    from gaussianmesh_b200._lib import lib
    ptrs = (C.c_void_p * 6)()
    lib.gm_geom_view(geom.data_ptr(), P, ptrs)
    base = geom.data_ptr()

    def view(k, dtype, shape):
        off = ptrs[k] - base
        cnt = int(np.prod(shape)) * torch.tensor([], dtype=dtype).element_size()
        return geom[off:off + cnt].view(dtype).view(*shape).clone()

    rgbc = view(4, torch.float32, (P, 4))
    return {"depths": view(0, torch.float32, (P,)), "means2D": view(1, torch.float32, (P, 2)),
            "cov3D": view(2, torch.float32, (P, 6)), "conic_opacity": view(3, torch.float32, (P, 4)),
            "rgb": rgbc[:, :3].contiguous(), "clamp_bits": rgbc[:, 3].contiguous().view(torch.int32),
            "tile_count": view(5, torch.int32, (num_tiles,))}


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max |b| -- the gradient parity measure (north star: 1e-3 rel)."""
    denom = float(b.abs().max())
    if denom == 0.0:
        return float(a.abs().max())
    return float((a - b).abs().max()) / denom


# Element-wise gradient measure ("norm-wise, plus element-wise with an absolute floor").  The max-norm
# measure above lets a gradient 100x below the tensor's largest entry be completely wrong; this one bounds every
# element: |a - b| <= rtol |b| + floor max|b|.  The reference's own gradients come out of float atomics in
# scheduling order, so a small fraction of heavily cancelling elements may differ between two runs of the reference
# itself; `min_frac` of the elements must pass and the worst offender is reported.
GRAD_RTOL = 1e-3
GRAD_FLOOR = 1e-5
GRAD_MIN_FRAC = 0.999


def grad_report(a: torch.Tensor, b: torch.Tensor, rtol: float = GRAD_RTOL, floor: float = GRAD_FLOOR) -> Dict[str, float]:
    a = a.detach().reshape(-1).double()
    b = b.detach().reshape(-1).double()
    scale = float(b.abs().max())
    diff = (a - b).abs()
    bound = rtol * b.abs() + floor * scale
    ok = diff <= bound
    excess = diff / bound.clamp_min(1e-300)
    worst = int(excess.argmax()) if a.numel() else 0
    return {"rel": float(diff.max()) / scale if scale > 0 else float(a.abs().max()) if a.numel() else 0.0,
            "frac_ok": float(ok.double().mean()) if a.numel() else 1.0,
            "worst": worst, "worst_got": float(a[worst]) if a.numel() else 0.0,
            "worst_want": float(b[worst]) if a.numel() else 0.0, "worst_excess": float(excess[worst]) if a.numel() else 0.0,
            "scale": scale}


def assert_grad(a: torch.Tensor, b: torch.Tensor, name: str, tol: float = 1e-3, min_frac: float = GRAD_MIN_FRAC) -> Dict[str, float]:
    """Both gradient gates: max-norm relative error <= tol AND the element-wise bound for >= min_frac of the elements."""
    rep = grad_report(a, b)
    assert rep["rel"] <= tol, f"grad {name}: max-norm rel err {rep['rel']:.3e} > {tol}"
    assert rep["frac_ok"] >= min_frac, (
        f"grad {name}: only {100 * rep['frac_ok']:.4f}% of the elements within {GRAD_RTOL}*|ref| + {GRAD_FLOOR}*max|ref|; "
        f"worst element {rep['worst']}: got {rep['worst_got']:.6e}, want {rep['worst_want']:.6e} "
        f"({rep['worst_excess']:.1f}x the bound, tensor scale {rep['scale']:.3e})")
    return rep
