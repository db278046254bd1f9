#!/usr/bin/env python
"""Where a view-parallel global step spends its time: CUDA-event sections of ViewParallelTrainer.step() per exchange mode,
per rank (rank 0 prints the mean / min / max over ranks).  Run under torch.distributed.run with N ranks.
Usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/vp_breakdown.py [--modes mc,p2p]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

from gaussianmesh_b200 import synthetic
from gaussianmesh_b200.cameras import upload_cameras
from gaussianmesh_b200.renderer import MeshGaussianModel
from gaussianmesh_b200.training import OptimizationParams
from gaussianmesh_b200.view_parallel import ViewParallelTrainer
from gaussianmesh_b200.view_shard import shard_views

W, H, P, NV = 1920, 1080, 1_000_000, 100


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--modes", default="mc,p2p,nccl")
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    Vm, Fm = synthetic.icosphere(4)
    arrays = synthetic.mesh_bound_scene(P, Vm, Fm, seed=0)
    cams_all = synthetic.orbit_cameras(NV, W, H)
    cams = upload_cameras([cams_all[i] for i in shard_views(NV, world, rank)], dev)
    bg = torch.zeros(3, device=dev)
    rng = np.random.default_rng(1)
    targets = [torch.from_numpy(rng.integers(0, 256, size=(3, H, W), dtype=np.uint8)).to(dev).float() / 255.0 for _ in range(2)]
    out = {}
    for mode in args.modes.split(","):
        model = MeshGaussianModel(arrays, dev, requires_grad=False)
        vp = ViewParallelTrainer(model, OptimizationParams(), W, H, mode=mode)
        vp.reserve_for(cams, bg)
        for i in range(5):
            vp.step(cams[i % len(cams)], bg, targets[i % 2])
        torch.cuda.synchronize()
        dist.barrier()
        vp.profile_sections = True
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            vp.step(cams[(5 + i) % len(cams)], bg, targets[i % 2])
        e1.record()
        sec = vp.section_times()
        sec["step_total"] = e0.elapsed_time(e1) / args.steps
        keys = sorted(sec)
        t = torch.tensor([sec[k] for k in keys], device=dev)
        allt = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        m = torch.stack(allt).cpu().numpy()
        out[mode] = {k: {"mean": round(float(m[:, i].mean()), 4), "min": round(float(m[:, i].min()), 4),
                         "max": round(float(m[:, i].max()), 4)} for i, k in enumerate(keys)}
        del vp, model
    if rank == 0:
        print(json.dumps({"world": world, "sections_ms": out}, indent=1))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
