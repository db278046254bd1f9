#!/usr/bin/env python
"""Lean timing of the full training iteration (bench section 5: 1M mesh-bound Gaussians on the 5,120-face proxy mesh, 1080p)
with the per-stage CUDA-event table, for kernel A/B work and ncu captures.
Usage: [GM_...=...] python scripts/time_iter.py [--steps K]   (prints one JSON line)"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from gaussianmesh_b200 import _lib, synthetic
from gaussianmesh_b200.cameras import upload_cameras
from gaussianmesh_b200.renderer import MeshGaussianModel
from gaussianmesh_b200.training import OptimizationParams, TrainingIteration

W, H, P, NV = 1920, 1080, 1_000_000, 100


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=30)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    V, F = synthetic.icosphere(4)
    arrays = synthetic.mesh_bound_scene(P, V, F, seed=0)
    cams = upload_cameras(synthetic.orbit_cameras(NV, W, H), dev)
    bg = torch.zeros(3, device=dev)
    rng = np.random.default_rng(1)
    targets = [torch.from_numpy(rng.integers(0, 256, size=(3, H, W), dtype=np.uint8)).to(dev).float() / 255.0 for _ in range(4)]
    it = TrainingIteration(MeshGaussianModel(arrays, dev, requires_grad=False), OptimizationParams(), W, H)
    it.reserve_for(cams, bg)

    def run(k, off=0):
        for i in range(k):
            it.step(cams[(off + i) % NV], bg, targets[i % 4])

    run(5)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run(args.steps, 5)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    _lib.profile_begin()
    run(args.steps, 5)
    torch.cuda.synchronize()
    prof = _lib.profile_end()
    assert not it.arena.verify()
    print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("GM_")}, "ms_per_iteration": round(ms, 4),
                      "stages": {k: round(t / n, 4) for k, (t, n) in prof.items()}, "counters": str(it.arena.last_info)}))


if __name__ == "__main__":
    main()
