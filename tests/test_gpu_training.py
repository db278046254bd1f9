"""GPU parity of the training-iteration pieces against oracle/train_np.py: photometric loss
(L1 + D-SSIM) with gradient, mesh-restrict loss, one-launch Adam, densification statistics, and the fused
TrainingIteration against the same iteration composed from autograd ops + the oracle optimizer.
Tolerances: loss values 1e-5 absolute (fp32 sums of ~1e5-1e6 terms, atomics order), image gradients 1e-3 relative
to the max norm like every other gradient of this repository (observed ~1e-5), Adam 1e-6 relative per step."""
import math

import numpy as np
import pytest
import torch

import scenes
from oracle import train_np

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape,lam", [((3, 37, 53), 0.2), ((3, 16, 16), 0.2), ((1, 5, 70), 1.0), ((3, 128, 200), 0.0),
                                       ((3, 11, 11), 0.5)])
def test_photometric_loss_matches_oracle(cuda_device, shape, lam):
    from gaussianmesh_b200.training import photometric_loss
    rng = np.random.default_rng(7)
    a = rng.uniform(0, 1, shape).astype(np.float32)
    b = np.clip(a + rng.normal(0, 0.15, shape), 0, 1).astype(np.float32)
    loss, l1, s, grad = train_np.photometric_loss_and_grad(a, b, lam)
    x = torch.from_numpy(a).to(cuda_device).requires_grad_(True)
    y = torch.from_numpy(b).to(cuda_device)
    got, parts = photometric_loss(x, y, lam, return_parts=True)
    (2.0 * got).backward()
    parts = parts.cpu().numpy()
    assert abs(float(got) - loss) <= 1e-5
    assert abs(parts[1] - l1) <= 1e-5 and abs(parts[2] - s) <= 1e-5
    err = np.abs(x.grad.cpu().numpy() / 2.0 - grad).max() / max(np.abs(grad).max(), 1e-12)
    assert err <= 1e-3, f"image gradient rel err {err:.3e}"


def test_ssim_and_l1_wrappers(cuda_device):
    from gaussianmesh_b200.training import ssim, l1_loss
    rng = np.random.default_rng(8)
    a = rng.uniform(0, 1, (3, 40, 40)).astype(np.float32)
    b = rng.uniform(0, 1, (3, 40, 40)).astype(np.float32)
    x = torch.from_numpy(a).to(cuda_device).requires_grad_(True)
    y = torch.from_numpy(b).to(cuda_device)
    s = ssim(x, y)
    s.backward()
    xo = torch.from_numpy(a).requires_grad_(True)
    so = train_np.ssim(xo, torch.from_numpy(b))
    so.backward()
    assert abs(float(s) - float(so)) <= 1e-5
    assert scenes.rel_err(x.grad.cpu(), xo.grad) <= 1e-3
    assert abs(float(ssim(y, y)) - 1.0) <= 1e-5
    assert abs(float(l1_loss(x, y)) - float(np.abs(a - b).mean())) <= 1e-6
    with pytest.raises(NotImplementedError):
        ssim(x, y, window_size=7)


def test_full_size_photometric_properties(cuda_device):
    """1920x1080: SSIM(x, x) = 1, loss(x, x) = 0; swapping a block of the image changes the gradient only within
    the 11x11 support around it (locality of the windowed statistics)."""
    from gaussianmesh_b200.training import photometric_loss
    g = torch.Generator(device="cpu").manual_seed(9)
    x = torch.rand(3, 1080, 1920, generator=g).to(cuda_device)
    y = torch.rand(3, 1080, 1920, generator=g).to(cuda_device)
    loss, parts = photometric_loss(x, x.clone(), 0.2, return_parts=True)
    assert abs(float(parts[2]) - 1.0) <= 1e-5 and abs(float(loss)) <= 1e-5
    xa = x.clone().requires_grad_(True)
    photometric_loss(xa, y, 0.2).backward()
    x2 = x.clone()
    x2[:, 500:516, 900:916] = 0.5
    xb = x2.requires_grad_(True)
    photometric_loss(xb, y, 0.2).backward()
    diff = (xa.grad - xb.grad).abs()
    outside = diff.clone()
    outside[:, 490:526, 890:926] = 0
    assert float(outside.max()) == 0.0 and float(diff.max()) > 0
    # crop check against the oracle on a window far from the borders' influence
    crop = (slice(None), slice(0, 64), slice(0, 96))
    _, _, _, go = train_np.photometric_loss_and_grad(x[crop].cpu().numpy(), y[crop].cpu().numpy(), 0.2)
    scale = (3 * 64 * 96) / x.numel()           # mean over the full image vs over the crop
    inner = (slice(None), slice(0, 50), slice(0, 80))
    err = np.abs(xa.grad[crop][inner].cpu().numpy() - go[inner] * scale).max() / np.abs(go * scale).max()
    assert err <= 1e-3


def test_mesh_restrict_loss_matches_oracle(cuda_device):
    from gaussianmesh_b200.training import mesh_restrict_loss
    rng = np.random.default_rng(10)
    P = 10_001
    scale = np.exp(rng.normal(-1.5, 1.0, (P, 3))).astype(np.float32)
    p1, p2, p3 = (rng.normal(0, 0.05, (P, 3)).astype(np.float32) for _ in range(3))
    loss, grad = train_np.mesh_restrict_loss(scale, p1, p2, p3, weight=6)
    t = lambda a: torch.from_numpy(a).to(cuda_device)
    s = t(scale).requires_grad_(True)
    got = mesh_restrict_loss(s, t(p1), t(p2), t(p3), weight=6)
    got.backward()
    assert abs(float(got) - loss) <= 1e-4 * max(1.0, loss)
    # Gaussians within rounding of the hinge may flip; everything else is exact
    margin = np.abs(scale.max(axis=1) - 6 * train_np.circumradius(p1, p2, p3)) > 1e-5
    assert np.array_equal(s.grad.cpu().numpy()[margin], grad[margin])
    assert grad.sum() > 100      # the case exercises the active branch


def test_adam_matches_oracle(cuda_device):
    from gaussianmesh_b200.training import Adam
    rng = np.random.default_rng(11)
    shapes = [(1000, 3), (1000, 1), (1000, 16, 3), (1000, 1), (1000, 3), (1000, 4), (7,), (4097,), (3, 5), (2,)]
    lrs = [1.6e-4, 1.6e-4, 0.0025 / 20, 0.05, 0.005, 0.001, 0.01, 0.02, 0.03, 0.04]
    params = [rng.normal(0, 1, s).astype(np.float32) for s in shapes]
    dev = [torch.from_numpy(p.copy()).to(cuda_device) for p in params]
    groups = [{"params": [d], "lr": lr, "name": str(i)} for i, (d, lr) in enumerate(zip(dev, lrs))]
    groups[2].update(lr_head=0.0025, period=48, split=3)
    opt = Adam(groups, lr=0.0, eps=1e-15)
    m = [np.zeros_like(p) for p in params]
    v = [np.zeros_like(p) for p in params]
    lr_arr = list(lrs)
    lr_arr[2] = np.where(np.arange(48) < 3, 0.0025, 0.0025 / 20).reshape(1, 16, 3)
    for n in range(1, 4):
        grads = [rng.normal(0, 1, s).astype(np.float32) * (rng.uniform(0, 1, s) > 0.3) for s in shapes]
        for d, g in zip(dev, grads):
            d.grad = torch.from_numpy(g.astype(np.float32)).to(cuda_device)
        opt.step()
        for i in range(len(params)):
            params[i], m[i], v[i] = train_np.adam_step(params[i], grads[i].astype(np.float32), m[i], v[i], lr_arr[i], n)
        for i, d in enumerate(dev):
            err = np.abs(d.cpu().numpy() - params[i]).max()
            assert err <= 2e-6, f"step {n} tensor {i}: {err:.3e}"
            assert np.abs(opt.state[id(d)]["exp_avg_sq"].cpu().numpy() - v[i]).max() <= 1e-6
    opt.zero_grad()
    assert all(d.grad is None for d in dev)


def test_densify_stats_match_oracle(cuda_device):
    from gaussianmesh_b200._lib import lib, check
    rng = np.random.default_rng(12)
    P = 5003
    radii = (rng.integers(0, 40, P) * (rng.uniform(0, 1, P) > 0.4)).astype(np.int32)
    grad = rng.normal(0, 1e-3, (P, 3)).astype(np.float32)
    mr, acc, den = (rng.uniform(0, 30, P).astype(np.float32), rng.uniform(0, 1, (P, 1)).astype(np.float32),
                    rng.integers(0, 5, (P, 1)).astype(np.float32))
    emr, eacc, eden = train_np.densify_stats(radii, grad, mr, acc, den)
    t = lambda a: torch.from_numpy(a.copy()).to(cuda_device)
    dr, dg, dmr, dacc, dden = t(radii), t(grad), t(mr), t(acc), t(den)
    check(lib.gm_densify_stats(P, dr.data_ptr(), dg.data_ptr(), dmr.data_ptr(), dacc.data_ptr(), dden.data_ptr(),
                               torch.cuda.current_stream().cuda_stream), "gm_densify_stats")
    assert np.array_equal(dmr.cpu().numpy(), emr) and np.array_equal(dden.cpu().numpy(), eden)
    assert np.abs(dacc.cpu().numpy() - eacc).max() <= 1e-7


@pytest.mark.parametrize("degree", [3, 1])
def test_training_iteration_matches_autograd_composition(cuda_device, degree):
    """Two iterations of TrainingIteration (fused, no tape) against the same iterations written the way
    train_mesh_gaussian.py:85-147 writes them: render() -> l1 / ssim / mesh_restrict_loss -> autograd backward ->
    densification statistics -> Adam, with the OPTIMIZER taken from the oracle."""
    from gaussianmesh_b200 import synthetic
    from gaussianmesh_b200.renderer import MeshGaussianModel, PipelineParams, render
    from gaussianmesh_b200.training import (OptimizationParams, TrainingIteration, mesh_restrict_loss, photometric_loss,
                                            get_expon_lr_func)
    dev = cuda_device
    P, W, H = 20_000, 320, 240
    V, F = synthetic.icosphere(3)
    arrays = synthetic.mesh_bound_scene(P, V, F, seed=2)
    opt = OptimizationParams(alpha_mrloss=0.05)              # many Gaussians larger than weight * sqrt(area): mrloss active
    fused_model = MeshGaussianModel(arrays, dev, sh_degree=degree, requires_grad=False)
    ref_model = MeshGaussianModel(arrays, dev, sh_degree=degree)
    it = TrainingIteration(fused_model, opt, W, H)
    cams = [scenes.camera(dev, W, H, index=i) for i in range(2)]
    bg = torch.zeros(3, device=dev)
    gts = [torch.rand(3, H, W, generator=torch.Generator().manual_seed(20 + i)).to(dev) for i in range(2)]

    names = ["_bc", "_distance", "_features", "_opacity", "_scaling", "_rotation"]
    lr_of = {"_features": np.where(np.arange(48) < 3, opt.feature_lr, opt.feature_lr / 20).reshape(1, 16, 3),
             "_opacity": opt.opacity_lr, "_scaling": opt.scaling_lr, "_rotation": opt.rotation_lr}
    sched = get_expon_lr_func(opt.position_lr_init, opt.position_lr_final, lr_delay_mult=opt.position_lr_delay_mult,
                              max_steps=opt.position_lr_max_steps)
    m = {k: np.zeros_like(getattr(ref_model, k).detach().cpu().numpy()) for k in names}
    v = {k: np.zeros_like(m[k]) for k in names}
    max_r = np.zeros(P, np.float32)
    acc = np.zeros((P, 1), np.float32)
    den = np.zeros((P, 1), np.float32)

    for n in (1, 2):
        # both paths start every iteration from the same parameters (Adam's +-lr steps would otherwise let them drift
        # apart by more than the integer radii tolerate); the oracle's moment estimates keep evolving on their own
        with torch.no_grad():
            for k in names:
                getattr(ref_model, k).copy_(getattr(fused_model, k))
        losses = it.step(cams[n - 1], bg, gts[n - 1]).clone()
        # --- the composed iteration
        for k in names:
            getattr(ref_model, k).grad = None
        ref_model.screenspace_points.grad = None
        pkg = render(cams[n - 1], ref_model, PipelineParams(), bg)
        photo, parts = photometric_loss(pkg["render"], gts[n - 1], opt.lambda_dssim, return_parts=True)
        mr = mesh_restrict_loss(pkg["scale"], pkg["vertex1"], pkg["vertex2"], pkg["vertex3"], weight=opt.alpha_mrloss)
        (photo + mr).backward()
        assert float(mr) > 0
        got = losses.cpu().numpy()
        assert abs(got[0] - float(photo)) <= 1e-5 and abs(got[3] - float(mr)) <= 1e-4 * float(mr)
        assert abs(got[1] - float(parts[1])) <= 1e-5 and abs(got[2] - float(parts[2])) <= 1e-5
        assert torch.equal(it.radii, pkg["radii"])
        for k, gk in [("_bc", "bc"), ("_distance", "distance"), ("_features", "sh"), ("_opacity", "opacity_logit"),
                      ("_scaling", "log_scale"), ("_rotation", "rot_raw")]:
            err = scenes.rel_err(it.grads[gk].view_as(getattr(ref_model, k)), getattr(ref_model, k).grad)
            assert err <= 1e-3, f"iteration {n} grad {k}: {err:.3e}"
        max_r, acc, den = train_np.densify_stats(pkg["radii"].cpu().numpy(), ref_model.screenspace_points.grad.cpu().numpy(),
                                                 max_r, acc, den)
        lr_of["_bc"] = lr_of["_distance"] = sched(n)
        with torch.no_grad():
            for k in names:
                p = getattr(ref_model, k)
                new, m[k], v[k] = train_np.adam_step(p.detach().cpu().numpy(), p.grad.cpu().numpy(), m[k], v[k], lr_of[k], n)
                p.copy_(torch.from_numpy(new).to(dev))
        for k in names:
            d = (getattr(fused_model, k).detach() - getattr(ref_model, k).detach()).abs()
            # an Adam step is ~ lr * g / |g| early on: elements whose gradient is ~0 in both paths may step in
            # different directions (<= 2 lr apart); everything else agrees to the gradient tolerance
            lr_max = float(np.max(lr_of[k]))
            assert float(d.max()) <= 2.002 * lr_max + 1e-7, f"iteration {n} param {k}"
            assert float((d > 2e-2 * lr_max).float().mean()) <= 0.02, f"iteration {n} param {k}"
    assert np.array_equal(it.max_radii2D.cpu().numpy(), max_r)
    assert np.array_equal(it.denom.cpu().numpy(), den)
    assert np.abs(it.bc_gradient_accum.cpu().numpy() - acc).max() <= 1e-3 * acc.max()


@pytest.mark.parametrize("N", [4, 5])
def test_densify_and_split_matches_oracle(cuda_device, N):
    """scene/mesh_based_gaussian_model.py:504-585: after two iterations (so the Adam moments and the statistics are
    populated), split every Gaussian above a threshold and compare every tensor of the model, the new mesh vertices /
    indices and the surviving optimizer state with the numpy restatement; then keep training on the new set."""
    from gaussianmesh_b200 import synthetic
    from gaussianmesh_b200.renderer import MeshGaussianModel
    from gaussianmesh_b200.training import OptimizationParams, TrainingIteration
    dev = cuda_device
    P, W, H = 6_000, 200, 136
    V, F = synthetic.icosphere(2)
    arrays = synthetic.mesh_bound_scene(P, V, F, seed=6)
    arrays["mesh_vertices"] = V.astype(np.float32)
    model = MeshGaussianModel(arrays, dev, requires_grad=False)
    it = TrainingIteration(model, OptimizationParams(), W, H)
    cam = scenes.camera(dev, W, H, index=0)
    bg = torch.zeros(3, device=dev)
    gt = torch.rand(3, H, W, generator=torch.Generator().manual_seed(70)).to(dev)
    for _ in range(2):
        it.step(cam, bg, gt)
    n = lambda t: t.detach().cpu().numpy().copy()
    grads = n(it.bc_gradient_accum / it.denom)
    grads[np.isnan(grads)] = 0.0
    thr = float(np.quantile(grads[grads > 0], 0.7))
    before = {"bc": n(model._bc), "distance": n(model._distance), "f_dc": n(model._features[:, :1]),
              "f_rest": n(model._features[:, 1:]), "opacity": n(model._opacity), "scaling": n(model._scaling),
              "rotation": n(model._rotation), "vertex1": n(model.vertex1), "vertex2": n(model.vertex2),
              "vertex3": n(model.vertex3), "normal": n(model.normal), "r": n(model.r), "fid": n(model.fid),
              "vertex_index": n(model.vertex_index), "v": n(model.v)}
    st = lambda attr: it.optimizer.state[id(getattr(model, attr))]
    feat_m, feat_v = n(st("_features")["exp_avg"]), n(st("_features")["exp_avg_sq"])
    moments = {"bc": (n(st("_bc")["exp_avg"]), n(st("_bc")["exp_avg_sq"])),
               "distance": (n(st("_distance")["exp_avg"]), n(st("_distance")["exp_avg_sq"])),
               "f_dc": (feat_m[:, :1], feat_v[:, :1]), "f_rest": (feat_m[:, 1:], feat_v[:, 1:]),
               "opacity": (n(st("_opacity")["exp_avg"]), n(st("_opacity")["exp_avg_sq"])),
               "scaling": (n(st("_scaling")["exp_avg"]), n(st("_scaling")["exp_avg_sq"])),
               "rotation": (n(st("_rotation")["exp_avg"]), n(st("_rotation")["exp_avg_sq"]))}
    want, want_mom, sel = train_np.densify_and_split(before, moments, grads, thr, N)
    S = it.densify_and_split(it.bc_gradient_accum / it.denom, thr, 0.0, N)
    assert S == int(sel.sum()) and 0 < S < P
    newP = P - S + S * N
    assert it.P == newP == model._bc.shape[0] and model.v.shape[0] == V.shape[0] + 3 * S
    eq = lambda a, b: np.array_equal(n(a), b)
    assert eq(model._bc, want["bc"]) and eq(model._distance, want["distance"]) and eq(model._opacity, want["opacity"])
    assert eq(model._rotation, want["rotation"]) and eq(model._features[:, :1], want["f_dc"]) and eq(model._features[:, 1:], want["f_rest"])
    assert np.abs(n(model._scaling) - want["scaling"]).max() <= 1e-6
    for k in ("vertex1", "vertex2", "vertex3", "normal", "r", "fid", "vertex_index", "v"):
        assert eq(getattr(model, k), want[k]), k
    fm, fv = n(st("_features")["exp_avg"]), n(st("_features")["exp_avg_sq"])
    assert np.array_equal(fm[:, :1], want_mom["f_dc"][0]) and np.array_equal(fv[:, 1:], want_mom["f_rest"][1])
    assert eq(st("_bc")["exp_avg"], want_mom["bc"][0]) and eq(st("_scaling")["exp_avg_sq"], want_mom["scaling"][1])
    assert float(it.denom.abs().max()) == 0.0 and it.max_radii2D.shape[0] == newP and it.optimizer.n_step == 2
    # children sit on their parent's face: the split triangles tile it
    area = lambda a, b, c: np.linalg.norm(np.cross(b - a, c - a), axis=1)
    kids = area(want["vertex1"][P - S:], want["vertex2"][P - S:], want["vertex3"][P - S:]).reshape(S, N)[:, :4].sum(axis=1)
    parents = area(before["vertex1"][sel], before["vertex2"][sel], before["vertex3"][sel])
    assert np.allclose(kids, parents, rtol=1e-4)
    # training continues on the new set
    losses = it.step(cam, bg, gt).cpu().numpy()
    assert np.isfinite(losses).all() and it.optimizer.n_step == 3
    # opacity reset (:334-339)
    want_op = train_np.reset_opacity(n(model._opacity))
    it.reset_opacity()
    assert np.abs(n(model._opacity) - want_op).max() <= 1e-5 and float(st("_opacity")["exp_avg"].abs().max()) == 0.0


def test_densification_schedule_inside_the_loop(cuda_device):
    """train_mesh_gaussian.py:114-139 through step(densify=True): statistics every iteration, densify_and_prune when
    iteration > densify_from_iter and iteration % densification_interval == 0 (that iteration skips the optimizer
    step, `update_flag`), opacity reset every opacity_reset_interval."""
    from gaussianmesh_b200 import synthetic
    from gaussianmesh_b200.renderer import MeshGaussianModel
    from gaussianmesh_b200.training import OptimizationParams, TrainingIteration
    dev = cuda_device
    P, W, H = 4_000, 160, 120
    V, F = synthetic.icosphere(2)
    arrays = synthetic.mesh_bound_scene(P, V, F, seed=8)
    arrays["mesh_vertices"] = V.astype(np.float32)
    opt = OptimizationParams(densify_from_iter=1, densification_interval=2, densify_grad_threshold=1e-7,
                             opacity_reset_interval=3)
    model = MeshGaussianModel(arrays, dev, requires_grad=False)
    it = TrainingIteration(model, opt, W, H)
    cam = scenes.camera(dev, W, H, index=0)
    bg = torch.zeros(3, device=dev)
    gt = torch.rand(3, H, W, generator=torch.Generator().manual_seed(80)).to(dev)
    it.step(cam, bg, gt, densify=True)                       # iteration 1: plain step
    assert it.P == P and it.optimizer.n_step == 1
    seen = int((it.denom > 0).sum())
    assert seen > 0
    it.step(cam, bg, gt, densify=True)                       # iteration 2: densify, no optimizer step
    assert it.optimizer.n_step == 1
    grown = it.P
    assert grown > P and (grown - P) % 4 == 0 and grown <= 5 * P   # the Gaussians seen so far split into five (N = 5, :126)
    assert model._bc.shape[0] == grown and float(it.denom.abs().max()) == 0.0
    it.step(cam, bg, gt, densify=True)                       # iteration 3: plain step + opacity reset
    assert it.optimizer.n_step == 2 and it.P == grown
    # reset to <= 0.01, then this iteration's Adam step (moments zeroed: at most one learning rate, 0.05, in the logit)
    assert float(model._opacity.max()) <= math.log(0.01 / 0.99) + opt.opacity_lr * 1.01
    assert np.isfinite(it.losses.cpu().numpy()).all()


def test_overflowed_frame_never_reaches_the_parameters(cuda_device):
    """An arena that is too small for a frame renders only the tiles that fit.  The sync-free iteration cannot see that on
    the host before the optimizer is enqueued, so the statistics and Adam are gated on the frame's overflow word ON THE DEVICE
    (gm_frame_overflow_flag, gm_adam_step_gated, gm_densify_stats_gated): the parameters and the statistics stay exactly as
    they were, the host counts the dropped iteration one step later, and the grown arena lets the next iteration through."""
    from gaussianmesh_b200 import synthetic
    from gaussianmesh_b200.renderer import MeshGaussianModel, upload_cameras
    from gaussianmesh_b200.training import OptimizationParams, TrainingIteration
    dev = cuda_device
    P, W, H = 8_000, 200, 136
    V, F = synthetic.icosphere(2)
    arrays = synthetic.mesh_bound_scene(P, V, F, seed=7)
    cams = upload_cameras(synthetic.orbit_cameras(3, W, H), dev)
    bg = torch.zeros(3, device=dev)
    gt = torch.rand(3, H, W, generator=torch.Generator().manual_seed(9)).to(dev)
    model = MeshGaussianModel(arrays, dev, requires_grad=False)
    it = TrainingIteration(model, OptimizationParams(alpha_mrloss=0.05), W, H)
    # an arena far too small, that will not be sized by the first frame
    it.arena._want = 64
    it.arena.headroom = 1.0
    before = {k: getattr(model, k).clone() for k in ("_features", "_bc", "_distance", "_scaling", "_rotation", "_opacity")}
    it.step(cams[0], bg, gt)
    torch.cuda.synchronize()
    need, visible, overflow, cap = it.arena._info[0].tolist()
    assert overflow == 1 and need > cap
    for k, v in before.items():
        assert torch.equal(getattr(model, k), v), k                       # the update was dropped on the device
    assert float(it.denom.sum()) == 0.0 and float(it.max_radii2D.max()) == 0.0
    it.step(cams[1], bg, gt)                                              # the arena has grown: this one goes through
    torch.cuda.synchronize()
    assert it.skipped_iterations == 1 and it.arena.overflowed == []
    assert not torch.equal(model._features, before["_features"])
    assert float(it.denom.sum()) > 0.0
    assert it.arena.verify() == []
