#!/usr/bin/env python
"""Timings of BASELINE.json configs 1 and 2 (the parity-test configurations, not bench lines) for BASELINE.md 2.3:
config 1 = 10K Gaussians, 256x256, forward only; config 2 = 100K mesh-bound Gaussians, 800x800, forward + L1 + backward.
Both arms on the same GPU: ours through TrainStep / ViewBatchRenderer, the reference through oracle/_ref driven like its glue.
Usage: python scripts/time_configs.py > gpurun_out/configs12.json"""
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import refcuda
from gaussianmesh_b200 import synthetic
from gaussianmesh_b200.cameras import upload_cameras
from gaussianmesh_b200.renderer import TrainStep, ViewBatchRenderer
from gaussianmesh_b200.mesh_gaussians import mesh_bind


def timed(fn, steps=100, warmup=10):
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(steps):
        fn(warmup + i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def main():
    dev = torch.device("cuda:0")
    out = {}
    bg = torch.zeros(3, device=dev)
    # ---- config 1
    P, W, H = 10_000, 256, 256
    a = synthetic.gaussian_scene(P, seed=0)
    sc = {k: torch.from_numpy(a[k]).to(dev) for k in ("means3D", "opacities", "shs", "scales", "rotations")}
    cams = upload_cameras(synthetic.orbit_cameras(8, W, H), dev)
    vb = ViewBatchRenderer(dev, sc["means3D"], sc["opacities"], shs=sc["shs"], scales=sc["scales"], rotations=sc["rotations"])
    vb.reserve_for(cams, bg)
    img = torch.empty(3, H, W, device=dev)
    ours = timed(lambda i: vb.render_into(cams[i % 8], bg, img))
    assert not vb.verify()

    def ref_fwd(i):
        c = cams[i % 8]
        return refcuda.RefFrame(bg, sc["means3D"], sc["opacities"], c.world_view_transform, c.full_proj_transform, c.camera_center,
                                math.tan(c.FoVx * 0.5), math.tan(c.FoVy * 0.5), H, W, 3, shs=sc["shs"], scales=sc["scales"],
                                rotations=sc["rotations"], sync=False)
    ref = timed(ref_fwd)
    fr = ref_fwd(0)
    vb.render_into(cams[0], bg, img)
    torch.cuda.synchronize()
    out["config1"] = {"workload": "10K Gaussians, 256x256, forward", "ours_ms": ours, "reference_ms": ref,
                      "max_abs_diff_vs_reference": float((img - fr.color).abs().max())}
    # ---- config 2
    P, W, H = 100_000, 800, 800
    V, F = synthetic.icosphere(4)
    m = synthetic.mesh_bound_scene(P, V, F, seed=0)
    t = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in m.items()}
    with torch.no_grad():
        xyz, scales, rot, opac = mesh_bind(t["bc_logits"], t["distance"], t["log_scales"], t["rot_raw"], t["opacity_logit"],
                                           t["vertex1"], t["vertex2"], t["vertex3"], t["normal"], t["r"])
    cams = upload_cameras(synthetic.orbit_cameras(8, W, H), dev)
    target = torch.rand(3, H, W, device=dev)
    ts = TrainStep(dev, xyz, opac, t["shs"], scales, rot, W, H)
    ts.reserve_for(cams, bg)
    ours = timed(lambda i: ts.step(cams[i % 8], bg, target))
    assert ts.verify() == 0

    def ref_train(i):
        c = cams[i % 8]
        fr = refcuda.RefFrame(bg, xyz, opac, c.world_view_transform, c.full_proj_transform, c.camera_center,
                              math.tan(c.FoVx * 0.5), math.tan(c.FoVy * 0.5), H, W, 3, shs=t["shs"], scales=scales, rotations=rot,
                              sync=False)
        diff = fr.color - target
        diff.abs().mean()
        return fr, fr.backward(torch.sign(diff) / diff.numel(), sync=False)
    ref = timed(ref_train)
    fr, rg = ref_train(0)
    ts.step(cams[0], bg, target)
    torch.cuda.synchronize()
    rel = float((ts.grads["means3D"].view(-1, 3) - rg["means3D"]).abs().max() / rg["means3D"].abs().max())
    out["config2"] = {"workload": "100K mesh-bound Gaussians (5,120-face proxy mesh), 800x800, forward + L1 + backward",
                      "ours_ms": ours, "reference_ms": ref, "grad_means3D_rel_err": rel,
                      "max_abs_diff_vs_reference": float((ts.image - fr.color).abs().max())}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
