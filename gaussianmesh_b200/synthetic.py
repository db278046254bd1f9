"""Seeded synthetic scenes, proxy meshes and cameras.

Host-side numpy only.  These generate the INPUTS of the hot path for tests and benchmarks -- the
reference ships no datasets and no network is available.  Camera matrices are built exactly as
the reference builds them (scene/cameras.py:42-51, utils/graphics_utils.py:38-71).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Tuple

import numpy as np

SH_C0 = 0.28209479177387814


# ---------------------------------------------------------------------------------------------
# cameras
# ---------------------------------------------------------------------------------------------
@dataclass
class Camera:
    """What `render()` reads from a reference `Camera` / `MiniCam` (scene/cameras.py)."""
    image_width: int
    image_height: int
    FoVx: float
    FoVy: float
    world_view_transform: np.ndarray   # [4,4] = W2C^T  (scene/cameras.py:48)
    full_proj_transform: np.ndarray    # [4,4] = W2C^T @ P^T  (scene/cameras.py:49-50)
    camera_center: np.ndarray          # [3]   (scene/cameras.py:51)

    @property
    def tanfovx(self) -> float:
        return math.tan(self.FoVx * 0.5)

    @property
    def tanfovy(self) -> float:
        return math.tan(self.FoVy * 0.5)

    def packed(self) -> np.ndarray:
        """[35] float32: viewmatrix(16) | projmatrix(16) | campos(3) -- the per-view host input."""
        return np.concatenate([self.world_view_transform.reshape(-1), self.full_proj_transform.reshape(-1),
                               self.camera_center.reshape(-1)]).astype(np.float32)


def get_world2view2(R: np.ndarray, t: np.ndarray, translate=np.array([0.0, 0.0, 0.0]), scale=1.0) -> np.ndarray:
    """utils/graphics_utils.py:38-49"""
    Rt = np.zeros((4, 4))
    Rt[:3, :3] = R.transpose()
    Rt[:3, 3] = t
    Rt[3, 3] = 1.0
    C2W = np.linalg.inv(Rt)
    cam_center = C2W[:3, 3]
    cam_center = (cam_center + translate) * scale
    C2W[:3, 3] = cam_center
    Rt = np.linalg.inv(C2W)
    return np.float32(Rt)


def get_projection_matrix(znear: float, zfar: float, fovX: float, fovY: float) -> np.ndarray:
    """utils/graphics_utils.py:51-71 (fp32 result, as jt.zeros(4,4) is float32)."""
    tanHalfFovY = math.tan(fovY / 2)
    tanHalfFovX = math.tan(fovX / 2)
    top = tanHalfFovY * znear
    bottom = -top
    right = tanHalfFovX * znear
    left = -right
    P = np.zeros((4, 4), dtype=np.float32)
    z_sign = 1.0
    P[0, 0] = 2.0 * znear / (right - left)
    P[1, 1] = 2.0 * znear / (top - bottom)
    P[0, 2] = (right + left) / (right - left)
    P[1, 2] = (top + bottom) / (top - bottom)
    P[3, 2] = z_sign
    P[2, 2] = z_sign * zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


def look_at_camera(eye, target, width: int, height: int, fovx_deg: float = 60.0, up=(0.0, 1.0, 0.0),
                   znear: float = 0.01, zfar: float = 100.0) -> Camera:
    """A camera at `eye` looking at `target`; COLMAP convention (x right, y down, z forward).
    R is camera-to-world (as the reference's Camera stores it), T = -R^T eye."""
    eye = np.asarray(eye, dtype=np.float64)
    target = np.asarray(target, dtype=np.float64)
    f = target - eye
    f = f / np.linalg.norm(f)
    upv = np.asarray(up, dtype=np.float64)
    r = np.cross(f, upv)
    r = r / np.linalg.norm(r)
    d = np.cross(f, r)          # "down"
    R = np.stack([r, d, f], axis=1)     # columns = camera axes in world coordinates
    T = -R.T @ eye
    fovx = math.radians(fovx_deg)
    fovy = 2.0 * math.atan(math.tan(fovx * 0.5) * height / width)
    wv = get_world2view2(R, T).transpose(1, 0)                               # cameras.py:48
    proj = get_projection_matrix(znear, zfar, fovx, fovy).transpose(1, 0)    # cameras.py:49
    full = (wv.astype(np.float32) @ proj.astype(np.float32)).astype(np.float32)   # cameras.py:50
    center = np.linalg.inv(wv)[3, :3].astype(np.float32)                     # cameras.py:51
    return Camera(width, height, fovx, fovy, np.ascontiguousarray(wv, dtype=np.float32), full, center)


def orbit_cameras(n: int, width: int, height: int, radius: float = 6.0, cam_height: float = 1.0,
                  fovx_deg: float = 60.0, phase: float = 0.0) -> List[Camera]:
    """n look-at-origin cameras evenly spaced on a circle in the xz-plane."""
    cams = []
    for i in range(n):
        a = phase + 2.0 * math.pi * i / max(n, 1)
        eye = (radius * math.sin(a), cam_height, radius * math.cos(a))
        cams.append(look_at_camera(eye, (0.0, 0.0, 0.0), width, height, fovx_deg))
    return cams


# ---------------------------------------------------------------------------------------------
# free Gaussians
# ---------------------------------------------------------------------------------------------
def gaussian_scene(P: int, seed: int = 0, extent: float = 2.0, log_scale_mean: float = math.log(0.01),
                   log_scale_std: float = 0.5, sh_rest_std: float = 0.05) -> Dict[str, np.ndarray]:
    """means ~ U([-extent,extent]^3), log-scales ~ N(log 0.01, 0.5^2), unit quaternions,
    opacity = sigmoid(N(0, 2^2)), f_dc = (U(0,1)-0.5)/C0, f_rest ~ N(0, 0.05^2); shs [P,16,3]."""
    rng = np.random.default_rng(seed)
    means = rng.uniform(-extent, extent, size=(P, 3)).astype(np.float32)
    log_scales = rng.normal(log_scale_mean, log_scale_std, size=(P, 3)).astype(np.float32)
    q = rng.normal(size=(P, 4))
    q = (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)
    opacity_logit = rng.normal(0.0, 2.0, size=(P, 1)).astype(np.float32)
    f_dc = ((rng.uniform(0.0, 1.0, size=(P, 1, 3)) - 0.5) / SH_C0).astype(np.float32)
    f_rest = rng.normal(0.0, sh_rest_std, size=(P, 15, 3)).astype(np.float32)
    return {
        "means3D": means,
        "log_scales": log_scales,
        "scales": np.exp(log_scales).astype(np.float32),
        "rotations": q,
        "opacity_logit": opacity_logit,
        "opacities": (1.0 / (1.0 + np.exp(-opacity_logit))).astype(np.float32),
        "shs": np.ascontiguousarray(np.concatenate([f_dc, f_rest], axis=1), dtype=np.float32),
    }


# ---------------------------------------------------------------------------------------------
# proxy mesh + mesh-bound Gaussians
# ---------------------------------------------------------------------------------------------
def icosphere(subdivisions: int = 4, radius: float = 1.5) -> Tuple[np.ndarray, np.ndarray]:
    """Icosphere: 4 subdivisions -> 2,562 vertices / 5,120 faces (the "5K-face proxy mesh")."""
    t = (1.0 + math.sqrt(5.0)) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    verts = [np.asarray(p, dtype=np.float64) / math.sqrt(1 + t * t) for p in v]
    faces = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2),
             (10, 7, 6), (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11),
             (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    for _ in range(subdivisions):
        cache: Dict[Tuple[int, int], int] = {}

        def mid(a: int, b: int) -> int:
            key = (a, b) if a < b else (b, a)
            if key not in cache:
                m = verts[a] + verts[b]
                verts.append(m / np.linalg.norm(m))
                cache[key] = len(verts) - 1
            return cache[key]

        nf = []
        for a, b, c in faces:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        faces = nf
    V = (np.asarray(verts) * radius).astype(np.float32)
    F = np.asarray(faces, dtype=np.int32)
    return V, F


def mesh_bound_scene(P: int, V: np.ndarray, F: np.ndarray, seed: int = 0) -> Dict[str, np.ndarray]:
    """Gaussian i is bound to face i mod |F| (parametrisation of
    scene/mesh_based_gaussian_model.py:139-152, per-Gaussian face data as in :205-219)."""
    rng = np.random.default_rng(seed)
    base = gaussian_scene(P, seed=seed + 1000)
    fid = (np.arange(P) % F.shape[0]).astype(np.int32)
    tri = F[fid]
    v1, v2, v3 = V[tri[:, 0]], V[tri[:, 1]], V[tri[:, 2]]
    n = np.cross(v2 - v1, v3 - v1)
    n = n / np.maximum(np.linalg.norm(n, axis=1, keepdims=True), 1e-20)
    a = np.linalg.norm(v1 - v2, axis=1, keepdims=True)
    b = np.linalg.norm(v2 - v3, axis=1, keepdims=True)
    c = np.linalg.norm(v3 - v1, axis=1, keepdims=True)
    out = dict(base)
    out.update({
        "face_id": fid,
        "triangles": np.ascontiguousarray(tri, dtype=np.int32),
        "vertex1": np.ascontiguousarray(v1, dtype=np.float32),
        "vertex2": np.ascontiguousarray(v2, dtype=np.float32),
        "vertex3": np.ascontiguousarray(v3, dtype=np.float32),
        "normal": n.astype(np.float32),
        "r": ((a + b + c) / 3.0).astype(np.float32),
        "bc_logits": rng.normal(size=(P, 3)).astype(np.float32),
        "distance": rng.normal(0.0, 0.5, size=(P, 1)).astype(np.float32),
        "rot_raw": (base["rotations"] * rng.uniform(0.5, 2.0, size=(P, 1))).astype(np.float32),
    })
    del out["means3D"]
    return out


def twist_bend_deformation(V: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Analytic deformation V' = Rot_y(0.6 y) V + 0.3 sin(pi x / 3) e_y and the per-vertex polar
    factors (R, S) of its deformation gradient (fp64 SVD on the host) -- the stand-in for
    pyACAP.GetRS (edittool/__init__.py:109), which is not installable.
    Returns (V' [Vn,3] f32, R [Vn,3,3] f32, S [Vn,3,3] f32) with F = R S."""
    x, y, z = V[:, 0].astype(np.float64), V[:, 1].astype(np.float64), V[:, 2].astype(np.float64)
    k = 0.6
    th = k * y
    c, s = np.cos(th), np.sin(th)
    Vd = np.stack([c * x + s * z, y + 0.3 * np.sin(math.pi * x / 3.0), -s * x + c * z], axis=1)
    Fg = np.zeros((V.shape[0], 3, 3))
    Fg[:, 0, 0] = c
    Fg[:, 0, 1] = k * (-s * x + c * z)
    Fg[:, 0, 2] = s
    Fg[:, 1, 0] = 0.3 * math.pi / 3.0 * np.cos(math.pi * x / 3.0)
    Fg[:, 1, 1] = 1.0
    Fg[:, 2, 0] = -s
    Fg[:, 2, 1] = k * (-c * x - s * z)
    Fg[:, 2, 2] = c
    U, sig, Vt = np.linalg.svd(Fg)
    R = U @ Vt
    S = np.swapaxes(Vt, 1, 2) @ (sig[:, :, None] * Vt)
    return Vd.astype(np.float32), R.astype(np.float32), S.astype(np.float32)

def packed_covariance(scales: np.ndarray, rotations: np.ndarray, modifier: float = 1.0) -> np.ndarray:
    """strip_symmetric(L L^T), L = R(normalize(q)) diag(modifier * s): build_covariance_from_scaling_rotation
    (utils/general_utils.py:64-109); fp64 internally, returned as [P,6] float32 (xx,xy,xz,yy,yz,zz)."""
    s = scales.astype(np.float64) * modifier
    q = rotations.astype(np.float64)
    q = q / np.linalg.norm(q, axis=1, keepdims=True)
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                  2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                  2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], axis=1).reshape(-1, 3, 3)
    L = R * s[:, None, :]
    S = L @ np.swapaxes(L, 1, 2)
    return np.stack([S[:, 0, 0], S[:, 0, 1], S[:, 0, 2], S[:, 1, 1], S[:, 1, 2], S[:, 2, 2]], axis=1).astype(np.float32)
