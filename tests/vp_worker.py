"""Worker of tests/test_gpu_multi.py: two ranks, one view each per step, view-parallel training in both exchange
modes against the same global steps computed by a single process (both views rendered sequentially, gradients
averaged on the host side of the test, replicated Adam).  Launched with torch.distributed.run."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch
import torch.distributed as dist


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    from gaussianmesh_b200 import synthetic
    from gaussianmesh_b200.renderer import MeshGaussianModel, upload_cameras
    from gaussianmesh_b200.training import OptimizationParams, TrainingIteration
    from gaussianmesh_b200.view_parallel import ViewParallelTrainer

    P, W, H, steps = 20_000, 320, 240, 2
    V, F = synthetic.icosphere(3)
    arrays = synthetic.mesh_bound_scene(P, V, F, seed=3)
    opt = OptimizationParams(alpha_mrloss=0.05)
    cams = upload_cameras(synthetic.orbit_cameras(world * steps, W, H), dev)
    bg = torch.zeros(3, device=dev)
    gts = [torch.rand(3, H, W, generator=torch.Generator().manual_seed(40 + i)).to(dev) for i in range(world * steps)]
    names = ["_features", "_bc", "_distance", "_scaling", "_rotation", "_opacity"]

    def run(mode):
        model = MeshGaussianModel(arrays, dev, requires_grad=False)
        tr = ViewParallelTrainer(model, opt, W, H, mode=mode)
        for s in range(steps):
            tr.step(cams[s * world + rank], bg, gts[s * world + rank])
        tr.merge_stats()              # the per-rank statistics are merged when they are consumed, not every step
        torch.cuda.synchronize()
        dist.barrier()
        before = ({k: getattr(model, k).detach().clone() for k in names},
                  (tr.it.max_radii2D.clone(), tr.it.bc_gradient_accum.clone(), tr.it.denom.clone()))
        # densification of the view-parallel model (every rank takes the same decisions), then one more global step
        split = tr.densify_and_prune(run.threshold, N=5)
        tr.step(cams[rank], bg, gts[rank])
        torch.cuda.synchronize()
        dist.barrier()
        after = {k: getattr(model, k).detach().clone() for k in names}
        return before[0], before[1], split, after

    def run_single():
        model = MeshGaussianModel(arrays, dev, requires_grad=False)
        it = TrainingIteration(model, opt, W, H)
        for s in range(steps):
            acc = torch.zeros_like(it.param_grads)
            for r in range(world):
                it.step(cams[s * world + r], bg, gts[s * world + r], iteration=s + 1, optimizer_step=False)
                acc += it.param_grads
            it.param_grads.copy_(acc / world)
            it.optimizer.step(it._grad_of)
        torch.cuda.synchronize()
        before = ({k: getattr(model, k).detach().clone() for k in names},
                  (it.max_radii2D.clone(), it.bc_gradient_accum.clone(), it.denom.clone()))
        g = (it.bc_gradient_accum / it.denom).nan_to_num(0.0).reshape(-1)
        thr = torch.quantile(g[g > 0], 0.97).reshape(1)
        dist.broadcast(thr, src=0)
        run.threshold = float(thr)
        split = it.densify_and_prune(run.threshold, N=5)
        acc = torch.zeros_like(it.param_grads)
        for r in range(world):
            it.step(cams[r], bg, gts[r], iteration=steps + 1, optimizer_step=False)
            acc += it.param_grads
        it.param_grads.copy_(acc / world)
        it.optimizer.step(it._grad_of)
        torch.cuda.synchronize()
        after = {k: getattr(model, k).detach().clone() for k in names}
        return before[0], before[1], split, after

    ref_p, ref_s, ref_split, ref_after = run_single()
    out = {"world": world}
    lr_max = {"_features": opt.feature_lr, "_bc": opt.position_lr_init, "_distance": opt.position_lr_init,
              "_scaling": opt.scaling_lr, "_rotation": opt.rotation_lr, "_opacity": opt.opacity_lr}
    for mode in ("nccl", "p2p", "mc"):
        try:
            got_p, got_s, got_split, got_after = run(mode)
        except Exception as ex:      # report, do not hang the other rank
            out[mode] = {"error": repr(ex)[:500]}
            continue
        res = {}
        for k in names:
            d = (got_p[k] - ref_p[k]).abs()
            res[k] = {"max_over_lr": float(d.max()) / lr_max[k], "frac_off": float((d > 2e-2 * lr_max[k]).float().mean())}
        res["max_radii_equal"] = bool(torch.equal(got_s[0], ref_s[0]))
        res["denom_equal"] = bool(torch.equal(got_s[2].view(-1), ref_s[2].view(-1)))
        res["accum_rel"] = float((got_s[1].view(-1) - ref_s[1].view(-1)).abs().max() / ref_s[1].abs().max())
        # parameters must be identical on every rank after a step
        flat = torch.cat([got_p[k].reshape(-1) for k in names])
        other = flat.clone()
        dist.broadcast(other, src=0)
        res["replicas_identical"] = bool(torch.equal(flat, other))
        # after densify_and_prune + one more step: same split, same shapes, parameters within an Adam step of the single
        # process, replicas still identical
        res["split"] = [int(got_split), int(ref_split)]
        res["after_shapes_equal"] = all(got_after[k].shape == ref_after[k].shape for k in names)
        if res["after_shapes_equal"]:
            res["after_max_over_lr"] = max(float((got_after[k] - ref_after[k]).abs().max()) / lr_max[k] for k in names)
            flat2 = torch.cat([got_after[k].reshape(-1) for k in names])
            other2 = flat2.clone()
            dist.broadcast(other2, src=0)
            res["after_replicas_identical"] = bool(torch.equal(flat2, other2))
        out[mode] = res
    gathered = [None] * world
    dist.all_gather_object(gathered, out)
    if rank == 0:
        print(json.dumps(gathered))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
