"""CPU tests of the oracle (oracle/cpu_rasterizer.c + oracle/python_path.py).

The reference ships no tests / golden vectors for this path, so the oracle is pinned against
  (1) tests/golden/*.npz -- outputs of the reference's own CUDA code recorded on a B200
      (tests/golden/make_golden.py through oracle/_ref), and
  (2) finite differences of its own float64 build (the backward formulas), and
  (3) independent formulations of the Python-side functions.
"""
import math
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

from gaussianmesh_b200 import synthetic  # noqa: E402


@pytest.fixture(scope="session")
def oracle():
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "cpu"], check=True, capture_output=True)
    from oracle import cpu_oracle
    return cpu_oracle


def _golden_cases():
    import make_golden
    return list(make_golden.CASES)


# ------------------------------------------------------------------------------------------------ (1) golden
@pytest.mark.parametrize("name", _golden_cases())
def test_oracle_matches_reference_cuda_golden(oracle, name):
    import make_golden
    arrays, cam, bg, degree, variant, colors, cov, dL, want_grads = make_golden.case_inputs(name)
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    gold = np.load(path)
    if colors is not None:
        assert np.array_equal(colors, gold["in_colors"])
    if cov is not None:
        assert np.array_equal(cov, gold["in_cov3D"])
    f = oracle.forward(arrays, cam, bg, degree, colors=colors, cov3D=cov)
    # integer / index results: exact
    assert np.array_equal(f["radii"], gold["radii"])
    assert f["num_rendered"] == int(gold["num_rendered"])
    assert np.array_equal(f["tiles_touched"], gold["state_tiles_touched"].astype(np.int32))
    assert np.array_equal(f["n_contrib"], gold["n_contrib"].astype(np.uint32))
    vis = gold["radii"] > 0
    if colors is None:
        assert np.array_equal(f["clamped"][vis].astype(bool), gold["state_clamped"][vis])
    # per-Gaussian floats: the CPU build does not contract into FMAs, so a few ulps of slack
    assert np.abs(f["depths"][vis] - gold["state_depths"][vis]).max() <= 2e-6
    assert np.abs(f["means2D"][vis] - gold["state_means2D"][vis]).max() <= 1e-4
    rel = np.abs(f["conic_opacity"][vis] - gold["state_conic_opacity"][vis]) / np.maximum(np.abs(gold["state_conic_opacity"][vis]), 1e-3)
    assert rel.max() <= 1e-2
    # image: 1e-4 everywhere except pixels where an alpha < 1/255 or T < 1e-4 decision flipped by rounding
    err = np.abs(f["color"] - gold["color"])
    assert (err > 1e-4).sum() <= max(3, err.size // 20000), f"{(err > 1e-4).sum()} pixels off"
    assert err.max() <= 1.0 / 255.0 + 1e-3
    assert np.abs(f["final_T"] - gold["final_T"]).max() <= 4e-3
    if want_grads:
        g = oracle.backward(arrays, cam, bg, degree, f, dL)
        names = {"means3D": "grad_means3D", "means2D": "grad_means2D", "opacities": "grad_opacity", "conic": "grad_conic",
                 "cov3D": "grad_cov3D"}
        if colors is None:
            names["shs"] = "grad_sh"
        else:
            names["colors"] = "grad_colors"
        if cov is None:
            names.update({"scales": "grad_scales", "rotations": "grad_rotations"})
        for k, gk in names.items():
            b = gold[gk]
            a = g[k].reshape(b.shape)
            rel = np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)
            assert rel <= 1e-3, f"{k}: {rel:.3e}"


# ------------------------------------------------------------------------------------------------ (2) finite differences
def _fd_scene(P=24, seed=3):
    arrays = synthetic.gaussian_scene(P, seed=seed, extent=0.8, log_scale_mean=math.log(0.25), log_scale_std=0.3)
    arrays["opacities"] = np.clip(arrays["opacities"], 0.2, 0.9)
    arrays = {k: v.astype(np.float64) for k, v in arrays.items()}
    cam = synthetic.orbit_cameras(5, 48, 40, radius=5.0)[1]
    return arrays, cam


def _loss(oracle, arrays, cam, bg, dL, degree, **kw):
    f = oracle.forward(arrays, cam, bg, degree, dtype=np.float64, **kw)
    return float((f["color"] * dL).sum()), f


@pytest.mark.parametrize("degree", [3, 1])
def test_backward_formulas_against_finite_differences(oracle, degree):
    arrays, cam = _fd_scene()
    bg = np.array([0.3, 0.1, 0.7])
    dL = np.random.default_rng(0).uniform(-1, 1, size=(3, cam.image_height, cam.image_width))
    _, f = _loss(oracle, arrays, cam, bg, dL, degree)
    assert (f["radii"] > 0).all()
    g = oracle.backward(arrays, cam, bg, degree, f, dL)
    rng = np.random.default_rng(1)
    for key, gkey in [("means3D", "means3D"), ("scales", "scales"), ("rotations", "rotations"), ("opacities", "opacities"),
                      ("shs", "shs")]:
        for trial in range(3):
            d = rng.normal(size=arrays[key].shape)
            if key == "shs" and degree < 3:
                d[:, (degree + 1) ** 2:, :] = 0.0
            eps = 1e-6
            plus = dict(arrays); plus[key] = arrays[key] + eps * d
            minus = dict(arrays); minus[key] = arrays[key] - eps * d
            fd = (_loss(oracle, plus, cam, bg, dL, degree)[0] - _loss(oracle, minus, cam, bg, dL, degree)[0]) / (2 * eps)
            an = float((g[gkey].reshape(d.shape) * d).sum())
            assert abs(fd - an) <= 2e-5 * max(1.0, abs(an)), f"{key} trial {trial}: fd {fd} vs analytic {an}"


def test_backward_precomputed_paths_against_finite_differences(oracle):
    arrays, cam = _fd_scene(seed=5)
    bg = np.array([1.0, 1.0, 1.0])
    dL = np.random.default_rng(2).uniform(-1, 1, size=(3, cam.image_height, cam.image_width))
    colors = np.random.default_rng(3).uniform(0, 1, size=(arrays["means3D"].shape[0], 3))
    cov = synthetic.packed_covariance(arrays["scales"], arrays["rotations"]).astype(np.float64)
    _, f = _loss(oracle, arrays, cam, bg, dL, 3, colors=colors, cov3D=cov)
    g = oracle.backward(arrays, cam, bg, 3, f, dL)
    rng = np.random.default_rng(4)
    eps = 1e-6
    d = rng.normal(size=colors.shape)
    fd = (_loss(oracle, arrays, cam, bg, dL, 3, colors=colors + eps * d, cov3D=cov)[0]
          - _loss(oracle, arrays, cam, bg, dL, 3, colors=colors - eps * d, cov3D=cov)[0]) / (2 * eps)
    an = float((g["colors"] * d).sum())
    assert abs(fd - an) <= 2e-5 * max(1.0, abs(an))
    d = rng.normal(size=cov.shape) * 1e-2
    fd = (_loss(oracle, arrays, cam, bg, dL, 3, colors=colors, cov3D=cov + eps * d)[0]
          - _loss(oracle, arrays, cam, bg, dL, 3, colors=colors, cov3D=cov - eps * d)[0]) / (2 * eps)
    an = float((g["cov3D"] * d).sum())     # packed off-diagonals carry both symmetric entries
    assert abs(fd - an) <= 2e-5 * max(1.0, abs(an))


def test_scale_gradient_is_with_respect_to_modified_scale(oracle):
    """backward.cu:298,318-320: dL_dscale is d/d(mod * s), NOT multiplied by mod -- a reference quirk kept."""
    arrays, cam = _fd_scene(seed=7)
    bg = np.zeros(3)
    dL = np.random.default_rng(5).uniform(-1, 1, size=(3, cam.image_height, cam.image_width))
    mod = 1.3
    _, f = _loss(oracle, arrays, cam, bg, dL, 3, scale_modifier=mod)
    g = oracle.backward(arrays, cam, bg, 3, f, dL)
    d = np.random.default_rng(6).normal(size=arrays["scales"].shape)
    eps = 1e-6
    plus = dict(arrays); plus["scales"] = arrays["scales"] + eps * d
    minus = dict(arrays); minus["scales"] = arrays["scales"] - eps * d
    fd = (_loss(oracle, plus, cam, bg, dL, 3, scale_modifier=mod)[0] - _loss(oracle, minus, cam, bg, dL, 3, scale_modifier=mod)[0]) / (2 * eps)
    an = float((g["scales"] * d).sum())
    assert abs(fd - mod * an) <= 2e-5 * max(1.0, abs(fd))


def test_float32_and_float64_builds_agree(oracle):
    arrays = synthetic.gaussian_scene(3000, seed=11, log_scale_mean=math.log(0.03))
    cam = synthetic.orbit_cameras(7, 160, 96)[4]
    bg = np.array([0.2, 0.4, 0.6], np.float32)
    f32 = oracle.forward(arrays, cam, bg, 3)
    f64 = oracle.forward(arrays, cam, bg, 3, dtype=np.float64)
    assert (f32["radii"] != f64["radii"]).sum() <= 2
    err = np.abs(f32["color"] - f64["color"])
    assert (err > 1e-4).sum() <= 5


# ------------------------------------------------------------------------------------------------ edge cases
def test_empty_scene_and_empty_tiles(oracle):
    cam = synthetic.orbit_cameras(3, 64, 48)[0]
    bg = np.array([0.25, 0.5, 0.75], np.float32)
    arrays = synthetic.gaussian_scene(0)
    f = oracle.forward(arrays, cam, bg, 3)
    assert f["num_rendered"] == 0 and f["radii"].size == 0
    assert np.array_equal(f["color"], np.broadcast_to(bg[:, None, None], (3, 48, 64)))   # ranges (0,0) -> background
    assert (f["final_T"] == 1).all() and (f["n_contrib"] == 0).all()
    g = oracle.backward(arrays, cam, bg, 3, f, np.ones((3, 48, 64), np.float32))
    assert g["means3D"].shape == (0, 3)


def test_near_plane_cull_and_mark_visible(oracle):
    cam = synthetic.orbit_cameras(4, 64, 64)[0]             # at (0,1,6) looking at the origin
    pts = np.array([[0, 0, 0], [0, 1, 6.5], [0, 1, 5.9], [0, 1, 5.79]], np.float32)
    vis = oracle.mark_visible(pts, cam)
    assert vis.tolist() == [True, False, False, True]       # view-space z must exceed 0.2 (auxiliary.h:153)
    arrays = synthetic.gaussian_scene(4, seed=1)
    arrays["means3D"] = pts
    f = oracle.forward(arrays, cam, np.zeros(3, np.float32), 3)
    assert ((f["radii"] > 0) <= vis).all() and f["radii"][1] == 0 and f["radii"][2] == 0


def test_sort_is_stable_on_equal_depth(oracle):
    """Two Gaussians at the same depth keep Gaussian-id order inside every tile (stable radix sort)."""
    cam = synthetic.orbit_cameras(4, 32, 32)[0]
    arrays = synthetic.gaussian_scene(6, seed=2, log_scale_mean=math.log(0.3))
    arrays["means3D"][:] = arrays["means3D"][0]
    f = oracle.forward(arrays, cam, np.zeros(3, np.float32), 3)
    longest = 0
    for lo, hi in f["ranges"]:
        ids = f["point_list"][lo:hi].tolist()
        assert ids == sorted(ids)
        longest = max(longest, len(ids))
    assert longest >= 4
    assert (np.diff(f["keys"].astype(np.uint64)) >= 0).all()


def test_ragged_image_size(oracle):
    arrays = synthetic.gaussian_scene(500, seed=3, log_scale_mean=math.log(0.05))
    cam = synthetic.orbit_cameras(4, 70, 37)[1]
    f = oracle.forward(arrays, cam, np.ones(3, np.float32), 2)
    assert f["color"].shape == (3, 37, 70) and np.isfinite(f["color"]).all()
    assert f["ranges"].shape == (5 * 3, 2)


# ------------------------------------------------------------------------------------------------ (3) python path
def test_python_covariance_matches_c_oracle_for_unit_quaternions():
    from oracle import python_path as pp
    arrays = synthetic.gaussian_scene(2000, seed=4)
    a = pp.build_covariance_from_scaling_rotation(arrays["scales"], 1.0, arrays["rotations"])
    b = synthetic.packed_covariance(arrays["scales"], arrays["rotations"])          # float64 closed form
    assert np.abs(a - b).max() <= 1e-6 * max(1.0, np.abs(b).max())
    # and it normalises the quaternion (utils/general_utils.py:78-80), unlike the CUDA path
    c = pp.build_covariance_from_scaling_rotation(arrays["scales"], 1.0, arrays["rotations"] * 3.0)
    assert np.abs(a - c).max() <= 1e-6


def test_python_sh_matches_c_oracle(oracle):
    from oracle import python_path as pp
    arrays = synthetic.gaussian_scene(1500, seed=6)
    cam = synthetic.orbit_cameras(4, 64, 64)[2]
    for deg in (0, 1, 2, 3):
        rgb = pp.sh_to_rgb(deg, arrays["shs"], arrays["means3D"], cam.camera_center)
        f = oracle.forward(arrays, cam, np.zeros(3, np.float32), deg)
        vis = f["radii"] > 0
        assert np.abs(rgb[vis] - f["rgb"][vis]).max() <= 2e-6


def test_python_mesh_bind_and_deform_identities():
    from oracle import python_path as pp
    V, F = synthetic.icosphere(2)
    arrays = synthetic.mesh_bound_scene(3000, V, F, seed=1)
    xyz = pp.get_xyz(arrays["bc_logits"], arrays["distance"], arrays["vertex1"], arrays["vertex2"], arrays["vertex3"],
                     arrays["normal"], arrays["r"])
    # with distance = 0 the point lies in the face plane and its barycentric weights are the softmax
    zero = np.zeros_like(arrays["distance"])
    proj = pp.get_xyz(arrays["bc_logits"], zero, arrays["vertex1"], arrays["vertex2"], arrays["vertex3"], arrays["normal"], arrays["r"])
    w = pp.get_barycentric_coordinate(proj.astype(np.float64), arrays["vertex1"].astype(np.float64),
                                      arrays["vertex2"].astype(np.float64), arrays["vertex3"].astype(np.float64))
    assert np.abs(w - pp.softmax(arrays["bc_logits"])).max() <= 1e-5
    off = ((xyz - proj) * arrays["normal"]).sum(axis=1, keepdims=True)
    assert np.abs(off - 4.0 * arrays["r"] * (pp.sigmoid(arrays["distance"]) - 0.5)).max() <= 1e-5
    # identity deformation leaves everything unchanged; a rigid rotation rotates positions and covariances
    Vn = V.shape[0]
    eye = np.broadcast_to(np.eye(3, dtype=np.float32), (Vn, 3, 3))
    cov6 = synthetic.packed_covariance(arrays["scales"], arrays["rotations"])
    cov = np.stack([cov6[:, 0], cov6[:, 1], cov6[:, 2], cov6[:, 1], cov6[:, 3], cov6[:, 4], cov6[:, 2], cov6[:, 4], cov6[:, 5]],
                   axis=1).reshape(-1, 3, 3)
    p2, c2, r2 = pp.deform_gaussian(V, V, eye, eye, arrays["triangles"], w, proj, cov)
    assert np.abs(p2 - proj).max() <= 1e-6 and np.abs(c2 - cov).max() <= 1e-7 and np.abs(r2 - eye[0]).max() <= 1e-6
    th = 0.7
    Q = np.array([[math.cos(th), 0, math.sin(th)], [0, 1, 0], [-math.sin(th), 0, math.cos(th)]], np.float32)
    # ACAP hands back R such that the local frame map is R^T (the reference transposes it, edittool/__init__.py:122)
    p3, c3, r3 = pp.deform_gaussian(V, V @ Q.T, np.broadcast_to(Q.T, (Vn, 3, 3)), eye, arrays["triangles"], w, proj, cov)
    assert np.abs(p3 - proj @ Q.T).max() <= 1e-5
    assert np.abs(c3 - Q @ cov @ Q.T).max() <= 1e-6 * max(1.0, np.abs(cov).max())
    assert np.abs(r3 - Q).max() <= 1e-6


# ------------------------------------------------------------------------------------------------ ACAP (8f-3)
def test_acap_oracle_matches_the_reference_golden_vectors():
    """oracle/acap_np.py against the vectors that ship inside the reference's ACAP zip (test/LOGRNEW.txt, test/S.txt for
    test/1.obj -> test/2.obj); the files carry 6 significant digits."""
    from oracle import acap_np
    g = np.load(os.path.join(ROOT, "tests", "golden", "acap_1_to_2.npz"))
    rest = acap_np.rest_state(g["V_rest"], g["F"])
    R, S = acap_np.get_rs(rest, g["V_deformed"])
    assert np.abs(R - g["R_gold"]).max() <= 3e-5 and np.abs(R - g["R_gold"]).mean() <= 2e-6
    assert np.abs(S - g["S_gold"]).max() <= 3e-5 and np.abs(S - g["S_gold"]).mean() <= 2e-6
    # polar factors: R^T (= r) is a proper rotation, S symmetric, r S reproduces the fitted affine map's action
    r = np.swapaxes(R, 1, 2)
    assert np.abs(r @ R - np.eye(3)).max() <= 1e-12 and np.all(np.linalg.det(r) > 0)
    assert np.abs(S - np.swapaxes(S, 1, 2)).max() <= 1e-12


def test_acap_identity_and_rigid_motion():
    from oracle import acap_np
    V, F = synthetic.icosphere(2)
    V = V.astype(np.float64)
    rest = acap_np.rest_state(V, F)
    R, S = acap_np.get_rs(rest, V)
    assert np.abs(R - np.eye(3)).max() <= 1e-12 and np.abs(S - np.eye(3)).max() <= 1e-12
    th = 0.9
    Q = np.array([[math.cos(th), -math.sin(th), 0], [math.sin(th), math.cos(th), 0], [0, 0, 1.0]])
    R, S = acap_np.get_rs(rest, 1.7 * V @ Q.T + np.array([0.3, -1.0, 2.0]))
    assert np.abs(np.swapaxes(R, 1, 2) - Q).max() <= 1e-10          # GetRS hands back r^T
    assert np.abs(S - 1.7 * np.eye(3)).max() <= 1e-10


def test_acap_ring_builder_matches_oracle():
    """The C++ one-ring builder of the library (host code, no GPU needed) against the Python restatement, on a closed
    mesh and on a mesh with boundary."""
    import ctypes as C
    from gaussianmesh_b200._lib import lib
    from oracle import acap_np
    V, F = synthetic.icosphere(2)
    open_F = F[: F.shape[0] // 2]                                       # half a sphere: boundary fans
    for faces in (F, open_F):
        faces = np.ascontiguousarray(faces, np.int32)
        Vn = V.shape[0]
        ro, rn = np.zeros(Vn + 1, np.int32), np.zeros(3 * len(faces) + Vn, np.int32)
        fo, fl = np.zeros(Vn + 1, np.int32), np.zeros(3 * len(faces), np.int32)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        assert lib.gm_acap_build_rings(Vn, len(faces), p(faces), p(ro), p(rn), p(fo), p(fl)) == 0
        off, nbr = acap_np.one_rings(Vn, faces)
        assert np.array_equal(ro, off)
        for v in range(Vn):
            a, b = rn[ro[v]:ro[v + 1]].tolist(), nbr[off[v]:off[v + 1]].tolist()
            assert sorted(a) == sorted(b)
            if a:       # same cyclic sequence (closed fans may start anywhere; chains start at the same end)
                k = b.index(a[0])
                assert a == b[k:] + b[:k]
            assert sorted(fl[fo[v]:fo[v + 1]].tolist()) == sorted(np.nonzero((faces == v).any(axis=1))[0].tolist())


# ------------------------------------------------------------------------------------------------ training pieces
def test_ssim_oracle_against_scipy_float64():
    """oracle/train_np.ssim (conv2d restatement of utils/loss_utils.py:36-82) against an independent float64
    separable-filter formulation (scipy.ndimage.correlate1d with zero padding)."""
    import torch
    from scipy.ndimage import correlate1d
    from oracle import train_np
    rng = np.random.default_rng(3)
    a = rng.uniform(0, 1, (3, 37, 53)).astype(np.float32)
    b = np.clip(a + rng.normal(0, 0.1, a.shape), 0, 1).astype(np.float32)
    w = np.exp(-(np.arange(11) - 5) ** 2 / (2 * 1.5 ** 2))
    w /= w.sum()
    filt = lambda x: correlate1d(correlate1d(x.astype(np.float64), w, axis=1, mode="constant"), w, axis=2, mode="constant")
    mu1, mu2 = filt(a), filt(b)
    s1, s2, s12 = filt(a * a.astype(np.float64)) - mu1 ** 2, filt(b * b.astype(np.float64)) - mu2 ** 2, filt(a.astype(np.float64) * b) - mu1 * mu2
    ref = (((2 * mu1 * mu2 + 1e-4) * (2 * s12 + 9e-4)) / ((mu1 ** 2 + mu2 ** 2 + 1e-4) * (s1 + s2 + 9e-4))).mean()
    got = float(train_np.ssim(torch.from_numpy(a), torch.from_numpy(b)))
    assert abs(got - ref) <= 2e-6
    assert abs(float(train_np.ssim(torch.from_numpy(a), torch.from_numpy(a))) - 1.0) <= 1e-6
    # window: normalised, symmetric, the 11 taps the device kernel hard-codes
    win = train_np.gaussian(11, 1.5).numpy()
    assert abs(win.sum() - 1) <= 1e-6 and np.allclose(win, win[::-1])
    assert np.allclose(win, [0.00102838, 0.00759876, 0.03600077, 0.10936069, 0.21300554, 0.26601172,
                             0.21300554, 0.10936069, 0.03600077, 0.00759876, 0.00102838], atol=1e-8)


def test_photometric_gradient_against_finite_differences():
    from oracle import train_np
    rng = np.random.default_rng(4)
    a = rng.uniform(0.1, 0.9, (3, 20, 23))
    b = rng.uniform(0.1, 0.9, (3, 20, 23))
    loss, l1, s, g = train_np.photometric_loss_and_grad(a, b, 0.2, dtype=__import__("torch").float64)
    assert abs(loss - (0.8 * l1 + 0.2 * (1 - s))) <= 1e-12
    for idx in [(0, 0, 0), (1, 10, 11), (2, 19, 22), (0, 5, 17)]:
        e = 1e-6
        ap, am = a.copy(), a.copy()
        ap[idx] += e
        am[idx] -= e
        fd = (train_np.photometric_loss_and_grad(ap, b, 0.2, dtype=__import__("torch").float64)[0]
              - train_np.photometric_loss_and_grad(am, b, 0.2, dtype=__import__("torch").float64)[0]) / (2 * e)
        assert abs(fd - g[idx]) <= 1e-7 + 1e-5 * abs(fd)


def test_adam_oracle_against_torch_adam():
    """jittor's Adam folds the bias correction into the step size and adds eps to the uncorrected sqrt(v);
    with eps = 1e-15 that is torch.optim.Adam to rounding."""
    import torch
    from oracle import train_np
    rng = np.random.default_rng(5)
    p0 = rng.normal(0, 1, 1000).astype(np.float32)
    tp = torch.nn.Parameter(torch.from_numpy(p0.copy()).double())
    opt = torch.optim.Adam([tp], lr=0.01, eps=1e-15)
    p, m, v = p0.copy(), np.zeros_like(p0), np.zeros_like(p0)
    for n in range(1, 6):
        g = rng.normal(0, 1, 1000).astype(np.float32)
        tp.grad = torch.from_numpy(g).double()
        opt.step()
        p, m, v = train_np.adam_step(p, g, m, v, 0.01, n)
    assert np.abs(p - tp.detach().numpy()).max() <= 2e-6
    # per-element learning rates (the f_dc / f_rest split of one feature tensor)
    lr = np.where(np.arange(96) % 48 < 3, 0.0025, 0.0025 / 20)
    q, _, _ = train_np.adam_step(np.zeros(96, np.float32), np.ones(96, np.float32), np.zeros(96, np.float32),
                                 np.zeros(96, np.float32), lr, 1)
    assert np.allclose(q, -lr, rtol=1e-5)


def test_mesh_restrict_lr_schedule_and_densify_stats_oracles():
    from oracle import train_np
    from gaussianmesh_b200.training import get_expon_lr_func
    # equilateral triangle of edge 2: |AB x AC| = 2 sqrt(3), "circumradius" = sqrt of that
    p1 = np.array([[0, 0, 0]], np.float32)
    p2 = np.array([[2, 0, 0]], np.float32)
    p3 = np.array([[1, math.sqrt(3), 0]], np.float32)
    r = math.sqrt(2 * math.sqrt(3))
    loss, g = train_np.mesh_restrict_loss(np.array([[0.1, 30.0, 0.2]], np.float32), p1, p2, p3, weight=10)
    assert abs(loss - (30.0 - 10 * r)) <= 1e-4 and g.tolist() == [[0.0, 1.0, 0.0]]
    loss, g = train_np.mesh_restrict_loss(np.array([[0.1, 3.0, 0.2]], np.float32), p1, p2, p3, weight=10)
    assert loss == 0.0 and not g.any()
    # schedule: host mirror == oracle, end points as documented (utils/general_utils.py:36-41)
    a = train_np.get_expon_lr_func(1.6e-4, 1.6e-6, lr_delay_mult=0.01, max_steps=30000)
    b = get_expon_lr_func(1.6e-4, 1.6e-6, lr_delay_mult=0.01, max_steps=30000)
    for step in (0, 1, 100, 15000, 30000, 40000):
        assert a(step) == b(step)
    assert abs(a(0) - 1.6e-4) <= 1e-12 and abs(a(30000) - 1.6e-6) <= 1e-12 and a(-1) == 0.0
    radii = np.array([0, 3, 7, 0], np.int32)
    gr = np.array([[3, 4, 9], [3, 4, 9], [0, 1, 9], [5, 5, 9]], np.float32)
    mr, acc, den = train_np.densify_stats(radii, gr, np.array([1, 5, 1, 9], np.float32), np.zeros((4, 1), np.float32),
                                          np.zeros((4, 1), np.float32))
    assert mr.tolist() == [1, 5, 7, 9] and acc[:, 0].tolist() == [0, 5, 1, 0] and den[:, 0].tolist() == [0, 1, 1, 0]
