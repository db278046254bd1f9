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
