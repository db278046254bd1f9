#!/usr/bin/env python
"""Lean timing of the headline training frame (1M Gaussians, 1080p, forward + L1 + backward) for kernel A/B work:
K steps without stage events, then K steps with every stage bracketed by CUDA events.  Kernel variants are chosen
through the library's environment switches, read once per process, hence one process per variant.
Usage: [GM_...=...] python scripts/time_step.py [--steps K] [--dump grads.npz]   (prints one JSON line)"""
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
from gaussianmesh_b200.renderer import TrainStep

W, H, P, NV = 1920, 1080, 1_000_000, 100


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--dump", default=None, help="write the gradients of view 0 to this .npz")
    ap.add_argument("--morton", action="store_true", help="experiment: the scene's Gaussians in 3-D Morton order")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    a = synthetic.gaussian_scene(P, seed=0)
    if args.morton:
        m = a["means3D"]
        qz = ((m - m.min(0)) / (m.max(0) - m.min(0) + 1e-9) * 1023).astype(np.uint64)

        def spread(v):
            v = (v | (v << 16)) & 0x030000FF
            v = (v | (v << 8)) & 0x0300F00F
            v = (v | (v << 4)) & 0x030C30C3
            return (v | (v << 2)) & 0x09249249
        order = np.argsort(spread(qz[:, 0]) | (spread(qz[:, 1]) << 1) | (spread(qz[:, 2]) << 2), kind="stable")
        a = {k: (np.ascontiguousarray(v[order]) if isinstance(v, np.ndarray) and v.shape[:1] == (P,) else v) for k, v in a.items()}
    sc = {k: torch.from_numpy(a[k]).to(dev) for k in ("means3D", "opacities", "shs", "scales", "rotations")}
    cams = upload_cameras(synthetic.orbit_cameras(NV, W, H), dev)
    bg = torch.zeros(3, device=dev)
    rng = np.random.default_rng(1)
    targets = [torch.from_numpy(rng.integers(0, 256, size=(3, H, W), dtype=np.uint8)).to(dev) for _ in range(4)]
    ts = TrainStep(dev, sc["means3D"], sc["opacities"], sc["shs"], sc["scales"], sc["rotations"], W, H)
    ts.reserve_for(cams, bg)

    def run(k, off=0):
        for i in range(k):
            ts.step(cams[(off + i) % NV], bg, targets[i % 4])

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
    assert ts.verify() == 0
    out = {"morton": args.morton, "env": {k: v for k, v in os.environ.items() if k.startswith("GM_")}, "ms_per_step": round(ms, 4),
           "stages": {k: round(t / n, 4) for k, (t, n) in prof.items()}}
    if args.dump:
        loss = ts.step(cams[0], bg, targets[0])
        torch.cuda.synchronize()
        g = {k: v.cpu().numpy() for k, v in ts.grads.items()}
        g["loss"] = np.asarray(float(loss))
        np.savez(args.dump, **g)
        out["dumped"] = sorted(g)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
