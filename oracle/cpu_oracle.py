"""TEST INFRASTRUCTURE ONLY -- numpy/ctypes front end of the CPU oracle (oracle/cpu_rasterizer.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; nothing under
gaussianmesh_b200/ does.  Build the libraries with `make -C oracle cpu`.

Pinning status: the reference has no tests or golden vectors for this path; the oracle is
pinned against outputs of the reference's own CUDA code recorded on a B200 (tests/golden/*.npz, see
tests/golden/make_golden.py) and against finite differences (tests/test_oracle.py).
"""
from __future__ import annotations

import ctypes as C
import math
import os
import time
from typing import Dict, Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}


def _lib(dtype):
    name = "libgmo_f32.so" if dtype == np.float32 else "libgmo_f64.so"
    if name not in _LIBS:
        path = os.path.join(HERE, "_build", name)
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle cpu`")
        l = C.CDLL(path)
        l.gmo_preprocess.restype = C.c_int64
        assert l.gmo_sizeof_real() == np.dtype(dtype).itemsize
        _LIBS[name] = l
    return _LIBS[name]


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _a(x, dtype) -> Optional[np.ndarray]:
    return None if x is None else np.ascontiguousarray(x, dtype=dtype)


def _view_scalars(cam, dtype):
    real = C.c_float if dtype == np.float32 else C.c_double
    return real(cam.tanfovx), real(cam.tanfovy)


def forward(arrays: Dict[str, np.ndarray], cam, bg, degree: int, colors: Optional[np.ndarray] = None,
            cov3D: Optional[np.ndarray] = None, scale_modifier: float = 1.0, dtype=np.float32, M: Optional[int] = None
            ) -> Dict[str, np.ndarray]:
    """The reference forward on the CPU.  `arrays` holds means3D, opacities and (shs | colors=...) and
    (scales, rotations | cov3D=...); `cam` is a gaussianmesh_b200.synthetic.Camera."""
    l = _lib(dtype)
    real = C.c_float if dtype == np.float32 else C.c_double
    means = _a(arrays["means3D"], dtype)
    P = means.shape[0]
    W, H = cam.image_width, cam.image_height
    shs = None if colors is not None else _a(arrays["shs"], dtype)
    colors = _a(colors, dtype)
    scales = None if cov3D is not None else _a(arrays["scales"], dtype)
    rots = None if cov3D is not None else _a(arrays["rotations"], dtype)
    cov3D = _a(cov3D, dtype)
    opac = _a(arrays["opacities"], dtype).reshape(-1)
    view = _a(cam.world_view_transform.reshape(-1), dtype)
    proj = _a(cam.full_proj_transform.reshape(-1), dtype)
    campos = _a(cam.camera_center, dtype)
    bg = _a(bg, dtype)
    if M is None:
        M = shs.shape[1] if shs is not None else 0
    st = {
        "radii": np.zeros(P, np.int32), "depths": np.zeros(P, dtype), "means2D": np.zeros((P, 2), dtype),
        "cov3D": np.zeros((P, 6), dtype), "conic_opacity": np.zeros((P, 4), dtype), "rgb": np.zeros((P, 3), dtype),
        "clamped": np.zeros((P, 3), np.uint8), "tiles_touched": np.zeros(P, np.int32),
    }
    tx, ty = _view_scalars(cam, dtype)
    R = l.gmo_preprocess(P, degree, M, W, H, tx, ty, real(scale_modifier), _p(view), _p(proj), _p(campos), _p(means),
                         _p(shs), _p(colors), _p(opac), _p(scales), _p(rots), _p(cov3D), _p(st["radii"]), _p(st["depths"]),
                         _p(st["means2D"]), _p(st["cov3D"]), _p(st["conic_opacity"]), _p(st["rgb"]), _p(st["clamped"]),
                         _p(st["tiles_touched"]))
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    keys = np.zeros(max(R, 1), np.uint64)
    ids = np.zeros(max(R, 1), np.uint32)
    ranges = np.zeros((tiles, 2), np.uint32)
    l.gmo_bin(P, W, H, _p(st["radii"]), _p(st["means2D"]), _p(st["depths"]), _p(st["tiles_touched"]), C.c_int64(R),
              _p(keys), _p(ids), _p(ranges))
    color = np.zeros((3, H, W), dtype)
    final_T = np.zeros((H, W), dtype)
    n_contrib = np.zeros((H, W), np.uint32)
    l.gmo_blend_forward(W, H, _p(ranges), _p(ids), _p(st["means2D"]), _p(st["conic_opacity"]), _p(st["rgb"]), _p(bg),
                        _p(color), _p(final_T), _p(n_contrib))
    if cov3D is not None:
        st["cov3D"] = cov3D
    st.update({"color": color, "final_T": final_T, "n_contrib": n_contrib, "num_rendered": int(R), "keys": keys[:R],
               "point_list": ids[:R], "ranges": ranges, "M": M, "dtype": dtype,
               "inputs": dict(means=means, shs=shs, colors=colors, scales=scales, rots=rots, cov3D=cov3D, opac=opac,
                              view=view, proj=proj, campos=campos, bg=bg, scale_modifier=scale_modifier)})
    return st


def backward(arrays, cam, bg, degree: int, fwd: Dict[str, np.ndarray], dL_dpix: np.ndarray) -> Dict[str, np.ndarray]:
    """The reference backward on the CPU, from the state `forward` returned.  Gradient layouts are the
    reference's (rasterize_points.py:302-310)."""
    dtype = fwd["dtype"]
    l = _lib(dtype)
    real = C.c_float if dtype == np.float32 else C.c_double
    i = fwd["inputs"]
    P = i["means"].shape[0]
    W, H = cam.image_width, cam.image_height
    M = fwd["M"]
    dL = _a(dL_dpix, dtype)
    g = {"means2D": np.zeros((P, 3), dtype), "conic": np.zeros((P, 4), dtype), "opacities": np.zeros((P, 1), dtype),
         "colors": np.zeros((P, 3), dtype), "means3D": np.zeros((P, 3), dtype), "cov3D": np.zeros((P, 6), dtype),
         "shs": np.zeros((P, max(M, 1), 3), dtype), "scales": np.zeros((P, 3), dtype), "rotations": np.zeros((P, 4), dtype)}
    R = fwd["num_rendered"]
    ids = np.ascontiguousarray(fwd["point_list"]) if R else np.zeros(1, np.uint32)
    l.gmo_blend_backward(P, W, H, _p(fwd["ranges"]), _p(ids), C.c_int64(R), _p(fwd["means2D"]), _p(fwd["conic_opacity"]),
                         _p(fwd["rgb"]), _p(i["bg"]), _p(fwd["final_T"]), _p(fwd["n_contrib"]), _p(dL), _p(g["means2D"]),
                         _p(g["conic"]), _p(g["opacities"]), _p(g["colors"]))
    tx, ty = _view_scalars(cam, dtype)
    cov_all = _a(fwd["cov3D"], dtype)
    l.gmo_geometry_backward(P, degree, M, W, H, tx, ty, real(i["scale_modifier"]), _p(i["view"]), _p(i["proj"]),
                            _p(i["campos"]), _p(i["means"]), _p(fwd["radii"]), _p(i["shs"]), _p(fwd["clamped"]),
                            _p(i["scales"]), _p(i["rots"]), _p(cov_all), _p(g["means2D"]), _p(g["conic"]), _p(g["colors"]),
                            _p(g["means3D"]), _p(g["cov3D"]), _p(g["shs"]), _p(g["scales"]), _p(g["rotations"]))
    if M == 0:
        g["shs"] = g["shs"][:, :0]
    return g


def mark_visible(means3D: np.ndarray, cam, dtype=np.float32) -> np.ndarray:
    means = _a(means3D, dtype)
    out = np.zeros(means.shape[0], np.uint8)
    _lib(dtype).gmo_mark_visible(means.shape[0], _p(means), _p(_a(cam.world_view_transform.reshape(-1), dtype)), _p(out))
    return out.astype(bool)


def timed_sample(P: int, W: int, H: int, frames: int = 1) -> dict:
    """bench.py's cpu_baseline: forward + L1 + backward of the benchmark workload on the host cores."""
    import sys
    sys.path.insert(0, os.path.dirname(HERE))
    from gaussianmesh_b200 import synthetic
    arrays = synthetic.gaussian_scene(P, seed=0)
    cams = synthetic.orbit_cameras(100, W, H)
    target = np.random.default_rng(1).uniform(0.0, 1.0, size=(3, H, W)).astype(np.float32)
    bg = np.zeros(3, np.float32)
    t0 = time.perf_counter()
    for f in range(frames):
        fwd = forward(arrays, cams[f % len(cams)], bg, 3)
        diff = fwd["color"] - target
        loss = float(np.abs(diff).mean())
        dL = (np.sign(diff) / diff.size).astype(np.float32)
        backward(arrays, cams[f % len(cams)], bg, 3, fwd, dL)
    dt = time.perf_counter() - t0
    cores = int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1))
    out = {"value": frames / dt, "unit": "frames/s", "cores": cores, "kind": "port",
           "sample": f"{frames} training frame(s) (forward + L1 + backward) of the same {P}-Gaussian {W}x{H} workload "
                     f"through oracle/cpu_rasterizer.c (C + OpenMP restatement of the reference), {dt:.2f} s, loss {loss:.6f}"}
    # the reference's Python (Jittor) preprocess chain restated in numpy fp32 (oracle/python_path.py): mesh bind,
    # activations, Python covariance and SH->RGB fallbacks, for the same number of Gaussians
    from . import python_path as pp
    V, F = synthetic.icosphere(4)
    mesh = synthetic.mesh_bound_scene(P, V, F, seed=0)
    t0 = time.perf_counter()
    inp = pp.mesh_bound_inputs(mesh)
    pp.build_covariance_from_scaling_rotation(inp["scales"], 1.0, inp["rotations"])
    pp.sh_to_rgb(3, inp["shs"], inp["means3D"], cams[0].camera_center)
    out["python_preprocess_ms"] = (time.perf_counter() - t0) * 1e3
    out["python_preprocess_note"] = ("numpy fp32 restatement of get_xyz/get_scaling/get_rotation/get_opacity + "
                                     "build_covariance_from_scaling_rotation + eval_sh for the same P (one frame), one thread")
    # the same chain as multi-threaded tensor ops (torch CPU standing in for Jittor's CPU backend)
    try:
        out.update(_torch_cpu_preprocess(mesh, cams[0].camera_center))
    except Exception as ex:      # a reported baseline, never a dependency
        out["jittor_cpu_preprocess_note"] = f"torch CPU restatement failed: {ex!r}"
    return out


def _torch_cpu_preprocess(mesh: dict, campos: np.ndarray) -> dict:
    """get_xyz / get_scaling / get_rotation / get_opacity (scene/mesh_based_gaussian_model.py:122-174), get_covariance
    (:24-29, utils/general_utils.py:64-109) and the convert_SHs_python colours (gaussian_renderer/__init__.py:87-92,
    utils/sh_utils.py:57-112) as elementwise tensor ops on the host cores, all threads."""
    import torch
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    t = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in mesh.items() if v.dtype == np.float32}
    cam = torch.from_numpy(np.ascontiguousarray(campos, dtype=np.float32))
    C0, C1 = 0.28209479177387814, 0.4886025119029199
    C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
    C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
          1.445305721320277, -0.5900435899266435]

    def chain():
        bc = torch.softmax(t["bc_logits"], dim=1)
        xyz = bc[:, 0:1] * t["vertex1"] + bc[:, 1:2] * t["vertex2"] + bc[:, 2:3] * t["vertex3"] \
            + 4.0 * t["r"] * (torch.sigmoid(t["distance"]) - 0.5) * t["normal"]
        scales = torch.exp(t["log_scales"])
        torch.sigmoid(t["opacity_logit"])
        q = t["rot_raw"] / torch.sqrt((t["rot_raw"] ** 2).sum(dim=1))[:, None]
        r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
        R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y), 2 * (x * y + r * z),
                         1 - 2 * (x * x + z * z), 2 * (y * z - r * x), 2 * (x * z - r * y), 2 * (y * z + r * x),
                         1 - 2 * (x * x + y * y)], dim=1).view(-1, 3, 3)
        L = R * scales[:, None, :]
        S = L @ L.transpose(1, 2)
        torch.stack([S[:, 0, 0], S[:, 0, 1], S[:, 0, 2], S[:, 1, 1], S[:, 1, 2], S[:, 2, 2]], dim=1)
        sh = t["shs"].transpose(1, 2)
        d = xyz - cam.repeat(xyz.shape[0], 1)
        d = d / d.norm(dim=1, keepdim=True)
        x, y, z = d[:, 0:1], d[:, 1:2], d[:, 2:3]
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        res = C0 * sh[..., 0] - C1 * y * sh[..., 1] + C1 * z * sh[..., 2] - C1 * x * sh[..., 3]
        res = (res + C2[0] * xy * sh[..., 4] + C2[1] * yz * sh[..., 5] + C2[2] * (2.0 * zz - xx - yy) * sh[..., 6]
               + C2[3] * xz * sh[..., 7] + C2[4] * (xx - yy) * sh[..., 8])
        res = (res + C3[0] * y * (3 * xx - yy) * sh[..., 9] + C3[1] * xy * z * sh[..., 10]
               + C3[2] * y * (4 * zz - xx - yy) * sh[..., 11] + C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[..., 12]
               + C3[4] * x * (4 * zz - xx - yy) * sh[..., 13] + C3[5] * z * (xx - yy) * sh[..., 14]
               + C3[6] * x * (xx - 3 * yy) * sh[..., 15])
        return torch.clamp(res + 0.5, min=0.0)

    chain()                                           # warm-up (thread pool, allocator)
    t0 = time.perf_counter()
    chain()
    ms = (time.perf_counter() - t0) * 1e3

    # the edit-time chain (a4-a6) for half as many Gaussians (config 5): deform_gaussian once per deformed mesh
    # (edittool/__init__.py:103-131), then per frame the rotated-direction eval_sh and strip_symmetric (:442-448,
    # edittool/general_utils.py:26-37) -- the reference recomputes strip_symmetric every frame
    Pe = t["shs"].shape[0] // 2
    g = torch.Generator().manual_seed(0)
    Vn = 2562
    V, Vd = torch.randn(Vn, 3, generator=g), torch.randn(Vn, 3, generator=g)
    VR, VS = torch.randn(Vn, 3, 3, generator=g), torch.randn(Vn, 3, 3, generator=g)
    tri = torch.randint(0, Vn, (Pe, 3), generator=g)
    w = torch.rand(Pe, 3, generator=g)
    pos = t["vertex1"][:Pe].clone()
    cov = torch.randn(Pe, 3, 3, generator=g)
    sh_e = t["shs"][:Pe]

    def deform():
        w_pos, w_rs = w[:, :, None], w[:, :, None, None]
        g_delta_pos = (w_pos * (Vd - V)[tri]).sum(dim=1)
        rot = (w_rs * VR[tri]).sum(dim=1).transpose(1, 2)
        g_delta_rs = torch.matmul(rot, (w_rs * VS[tri]).sum(dim=1))
        return pos + g_delta_pos, torch.matmul(torch.matmul(g_delta_rs, cov), g_delta_rs.transpose(1, 2)), rot

    def edit_frame(dpos, dcov, rot):
        sh = sh_e.transpose(1, 2)
        d = dpos - cam.repeat(dpos.shape[0], 1)
        d = d / d.norm(dim=1, keepdim=True)
        d = torch.matmul(rot.transpose(1, 2), d.unsqueeze(2)).squeeze(2)
        x, y, z = d[:, 0:1], d[:, 1:2], d[:, 2:3]
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        res = C0 * sh[..., 0] - C1 * y * sh[..., 1] + C1 * z * sh[..., 2] - C1 * x * sh[..., 3]
        res = (res + C2[0] * xy * sh[..., 4] + C2[1] * yz * sh[..., 5] + C2[2] * (2.0 * zz - xx - yy) * sh[..., 6]
               + C2[3] * xz * sh[..., 7] + C2[4] * (xx - yy) * sh[..., 8])
        res = (res + C3[0] * y * (3 * xx - yy) * sh[..., 9] + C3[1] * xy * z * sh[..., 10]
               + C3[2] * y * (4 * zz - xx - yy) * sh[..., 11] + C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[..., 12]
               + C3[4] * x * (4 * zz - xx - yy) * sh[..., 13] + C3[5] * z * (xx - yy) * sh[..., 14]
               + C3[6] * x * (xx - 3 * yy) * sh[..., 15])
        torch.clamp(res + 0.5, min=0.0)
        c6 = torch.zeros(dcov.shape[0], 6)
        c6[:, 0], c6[:, 1], c6[:, 2] = dcov[:, 0, 0], dcov[:, 0, 1], dcov[:, 0, 2]
        c6[:, 3], c6[:, 4], c6[:, 5] = dcov[:, 1, 1], dcov[:, 1, 2], dcov[:, 2, 2]

    state = deform()
    t0 = time.perf_counter()
    state = deform()
    deform_ms = (time.perf_counter() - t0) * 1e3
    edit_frame(*state)
    t0 = time.perf_counter()
    edit_frame(*state)
    edit_ms = (time.perf_counter() - t0) * 1e3
    return {"jittor_cpu_preprocess_ms": ms, "jittor_cpu_preprocess_cores": threads,
            "jittor_cpu_edit_frame_ms": edit_ms, "jittor_cpu_deform_ms": deform_ms, "jittor_cpu_edit_gaussians": int(Pe),
            "jittor_cpu_preprocess_note": "the reference's Python preprocess chain (bind + activations + Python covariance + "
                                          "Python SH colours) as multi-threaded tensor ops, torch CPU standing in for Jittor's "
                                          "CPU backend (not installable here), one frame, same P"}
