#!/usr/bin/env python
"""Benchmark of the hot path: 1M Gaussians @ 1920x1080, forward + L1 + full backward ("training frame"),
independent views sharded over N GPUs (one process per GPU, no collective on the render path).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Rank 0 prints ONE JSON line.  Keys (see DESIGN.md "Measurement"):
  value / ms_per_step   training frames per second over all ranks, inputs resident in HBM, CUDA-event timed,
                        max over ranks
  e2e                   the same step driven with HOST inputs: per step the camera (140 B) and the 8-bit target image
                        (6.2 MB, as the reference's loader holds it before / 255.0) are copied from pinned host memory and
                        the loss is read back
  forward               forward-only rendering of the rank's view shard (config 3), frames/s
  edit                  config 5: deform once, then rotated-direction SH colours + forward per orbit frame
  train_iteration       the whole iteration of train_mesh_gaussian.py:73-147 on mesh-bound Gaussians (bind, render,
                        L1 + D-SSIM + mesh-restrict loss, backward, densification statistics, Adam), iterations/s
  roofline              the dominant kernel: algorithmic bytes / CUDA-event time / measured HBM peak
  cpu_baseline          the CPU oracle (oracle/, C + OpenMP restatement of the reference) on the host cores,
                        bounded sample, rank 0 at N=1 only
`--impl reference` times the UNMODIFIED reference CUDA rasterizer (oracle/_ref, rebuilt for sm_100a) driven the
way its Jittor glue drives it, on the same workload, every rank on its own view shard.
"""
from __future__ import annotations

import argparse
import gc
import json
import math
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch

P_GAUSS = 1_000_000
WIDTH, HEIGHT = 1920, 1080
NUM_VIEWS = 100
TARGET_POOL = 4
FORWARD_LANES = int(os.environ.get("GM_FORWARD_LANES", "2"))    # frames in flight in the forward-only section
METRIC = "train_frames_per_sec_1M_gaussians_1080p"
UNIT = "frames/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--gaussians", type=int, default=P_GAUSS, help="debug only: a smaller scene is NOT the benchmark")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-presize", action="store_true", help="profiling runs: skip the arena sizing pass over the views")
    ap.add_argument("--cpu-sample-frames", type=int, default=1)
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi sampled every 200 ms DURING the timed regions (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.tmp,
                                         stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.tmp.flush()
        self.tmp.seek(0)
        sm, mx, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.tmp.read().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.tmp.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # "under load" = samples in the upper half of the observed power range
        thr = (max(power) + min(power)) / 2
        loaded = [s for s, p in zip(sm, power) if p >= thr] or sm
        return {"sm_mhz": statistics.median(loaded), "sm_max_mhz": max(mx), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


def nvlink_bytes(gpu_index: int):
    """(tx, rx) data bytes over all NVLink links of one GPU since driver load (`nvidia-smi nvlink -gt d`), or None."""
    import re
    try:
        out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(gpu_index)], capture_output=True, text=True,
                             timeout=30).stdout
    except (OSError, subprocess.SubprocessError):
        return None
    tx = [int(m) for m in re.findall(r"Data Tx:\s*(\d+)\s*KiB", out)]
    rx = [int(m) for m in re.findall(r"Data Rx:\s*(\d+)\s*KiB", out)]
    if not tx or not rx:
        return None
    return sum(tx) * 1024, sum(rx) * 1024


# ------------------------------------------------------------------------------------------------ workload
def build_workload(device, P, rank, world):
    from gaussianmesh_b200 import synthetic
    from gaussianmesh_b200.cameras import upload_cameras
    from gaussianmesh_b200.view_shard import shard_views
    arrays = synthetic.gaussian_scene(P, seed=0)
    scene = {k: torch.from_numpy(arrays[k]).to(device) for k in ("means3D", "opacities", "shs", "scales", "rotations")}
    cams_host = synthetic.orbit_cameras(NUM_VIEWS, WIDTH, HEIGHT)
    mine = list(shard_views(NUM_VIEWS, world, rank))
    cams_host = [cams_host[i] for i in mine]
    cams = upload_cameras(cams_host, device)
    # targets are 8-bit images, as the reference's loader reads them (float = uint8 / 255.0, utils/general_utils.py:22-27):
    # the host copies stay uint8 (6.2 MB each), the resident copies are uint8 as well as float32
    rng = np.random.default_rng(1)
    targets_host = [torch.from_numpy(rng.integers(0, 256, size=(3, HEIGHT, WIDTH), dtype=np.uint8)).pin_memory()
                    for _ in range(TARGET_POOL)]
    targets_u8 = [t.to(device) for t in targets_host]
    targets = [(t.double() / 255.0).float() for t in targets_u8]      # IEEE uint8 / 255.0 in fp32 (torch.div by a scalar multiplies)
    build_workload.targets_u8 = targets_u8
    cams_packed_host = torch.from_numpy(np.stack([c.packed() for c in cams_host])).pin_memory()
    return scene, cams_host, cams, targets_host, targets, cams_packed_host


class RefHyper:
    """arguments/__init__.py:71-91 defaults used by the reference-style iteration (kept here so that the reference arm
    never imports gaussianmesh_b200.training, which maps libCudaRasterizer.so)."""
    position_lr_init = 0.00016
    feature_lr = 0.0025
    opacity_lr = 0.05
    scaling_lr = 0.005
    rotation_lr = 0.001
    lambda_dssim = 0.2
    alpha_mrloss = 6


EDIT_P, EDIT_VIEWS = 500_000, 200


class RefEditObject:
    """Config-5 object for the reference arm, built with tensor ops only (torch standing in for Jittor): bind on the
    face, numpy barycentric weights (the reference's get_barycentric_coordinate), analytic per-vertex R / S (pyACAP is a
    CPU library of the reference that cannot be installed), deform_gaussian as its op chain."""

    def __init__(self, device, rank, world):
        import refcuda
        from gaussianmesh_b200 import synthetic
        from gaussianmesh_b200.cameras import upload_cameras
        from gaussianmesh_b200.view_shard import shard_views
        from oracle import python_path
        V, F = synthetic.icosphere(4)
        a = synthetic.mesh_bound_scene(EDIT_P, V, F, seed=0)
        t = {k: torch.from_numpy(np.ascontiguousarray(v)).to(device) for k, v in a.items()}
        bc = torch.softmax(t["bc_logits"], dim=1)
        pos = (bc[:, 0:1] * t["vertex1"] + bc[:, 1:2] * t["vertex2"] + bc[:, 2:3] * t["vertex3"]).contiguous()
        scales = torch.exp(t["log_scales"])
        rots = torch.nn.functional.normalize(t["rot_raw"], dim=1)
        self.opacity = torch.sigmoid(t["opacity_logit"]).contiguous()
        self.shs = t["shs"]
        Vd64 = V.astype(np.float64)
        tri = a["triangles"]
        w = python_path.get_barycentric_coordinate(pos.cpu().numpy().astype(np.float64), Vd64[tri[:, 0]], Vd64[tri[:, 1]],
                                                   Vd64[tri[:, 2]]).astype(np.float32)
        cov6 = torch.from_numpy(synthetic.packed_covariance(scales.cpu().numpy(), rots.cpu().numpy())).to(device)
        Vd, R, S = synthetic.twist_bend_deformation(V)
        f = lambda x: torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(device)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        args = (f(V), f(Vd), f(R), f(S), torch.from_numpy(tri).to(device), f(w), pos, cov6)
        refcuda.deform_gaussians_torch(*args)
        torch.cuda.synchronize()
        ev0.record()
        self.deform_pos, self.deform_cov6, self.deform_rot = refcuda.deform_gaussians_torch(*args)
        ev1.record()
        torch.cuda.synchronize()
        self.deform_ms = ev0.elapsed_time(ev1)
        cams_host = synthetic.orbit_cameras(EDIT_VIEWS, WIDTH, HEIGHT)
        self.cams = upload_cameras([cams_host[i] for i in shard_views(EDIT_VIEWS, world, rank)], device)


def build_edit_workload(device, rank, world):
    """Config 5: 500K mesh-bound Gaussians on the 5,120-face proxy mesh, rest pose -> twist-and-bend deformation,
    200-frame orbit sharded over ranks."""
    from gaussianmesh_b200 import synthetic
    from gaussianmesh_b200.renderer import DeformedObject, shard_views, upload_cameras
    from gaussianmesh_b200.mesh_gaussians import mesh_bind
    V, F = synthetic.icosphere(4)
    a = synthetic.mesh_bound_scene(EDIT_P, V, F, seed=0)
    t = {k: torch.from_numpy(np.ascontiguousarray(v)).to(device) for k, v in a.items()}
    with torch.no_grad():
        pos, scales, rots, opac = mesh_bind(t["bc_logits"], torch.zeros_like(t["distance"]), t["log_scales"], t["rot_raw"],
                                            t["opacity_logit"], t["vertex1"], t["vertex2"], t["vertex3"], t["normal"], t["r"])
    cov6 = torch.from_numpy(synthetic.packed_covariance(scales.cpu().numpy(), rots.cpu().numpy())).to(device)
    obj = DeformedObject.load_mesh(pos, cov6, opac, t["shs"], pos, a["face_id"], V.astype(np.float64), F, device)
    Vd, _, _ = synthetic.twist_bend_deformation(V)
    # per-vertex rotation / shear of the deformation: ACAP GetRS on the GPU (the reference calls pyACAP on the CPU)
    from gaussianmesh_b200.acap import pyACAP
    tool = pyACAP((V, F), device=device)
    tool.GetRS(V, Vd, 1, 0)                                    # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    R1, S1 = tool.GetRS(V, Vd, 1, 0)
    torch.cuda.synchronize()
    acap_ms = (time.perf_counter() - t0) * 1e3
    cams_host = synthetic.orbit_cameras(EDIT_VIEWS, WIDTH, HEIGHT)
    cams_host = [cams_host[i] for i in shard_views(EDIT_VIEWS, world, rank)]
    return obj, (Vd, R1.reshape(-1, 3, 3), S1.reshape(-1, 3, 3)), upload_cameras(cams_host, device), acap_ms


def build_iteration_workload(P):
    """Section 5: BASELINE config 2 at the headline size -- P mesh-bound Gaussians on the 5,120-face proxy mesh."""
    from gaussianmesh_b200 import synthetic
    V, F = synthetic.icosphere(4)
    return synthetic.mesh_bound_scene(P, V, F, seed=0)


class OursIteration:
    """train_mesh_gaussian.py:73-147 through gaussianmesh_b200.training.TrainingIteration."""

    def __init__(self, device, arrays, W, H):
        from gaussianmesh_b200.renderer import MeshGaussianModel
        from gaussianmesh_b200.training import OptimizationParams, TrainingIteration
        self.model = MeshGaussianModel(arrays, device, requires_grad=False)
        self.it = TrainingIteration(self.model, OptimizationParams(), W, H)

    def setup(self, cams, bg):
        self.it.reserve_for(cams, bg)

    def step(self, cam, bg, gt):
        return self.it.step(cam, bg, gt)

    def check(self):
        if self.it.arena.verify():
            raise RuntimeError("arena overflow inside the timed training-iteration region")

    def counters(self):
        return self.it.arena.last_info


class ReferenceIteration:
    """The same iteration the way the reference writes it, torch standing in for Jittor (which cannot be installed):
    elementwise bind / activations with an autograd tape (scene/mesh_based_gaussian_model.py:122-174, incl. the
    per-frame concat of f_dc / f_rest), the UNMODIFIED reference CUDA rasterizer (oracle/_ref), conv2d SSIM + L1 +
    mesh_restrict_loss (utils/loss_utils.py, restated in oracle/train_np.py), indexed densification statistics
    (train_mesh_gaussian.py:117-121) and Adam over the seven parameter groups."""

    def __init__(self, device, arrays, W, H):
        import refcuda
        self.rc, self.W, self.H = refcuda, W, H
        t = {k: torch.from_numpy(np.ascontiguousarray(v)).to(device) for k, v in arrays.items()}
        self.t = t
        leaf = lambda x: x.clone().requires_grad_(True)
        self.bc, self.distance = leaf(t["bc_logits"]), leaf(t["distance"])
        self.f_dc, self.f_rest = leaf(t["shs"][:, :1].contiguous()), leaf(t["shs"][:, 1:].contiguous())
        self.opacity, self.scaling, self.rotation = leaf(t["opacity_logit"]), leaf(t["log_scales"]), leaf(t["rot_raw"])
        o = self.o = RefHyper
        self.optimizer = torch.optim.Adam([
            {"params": [self.bc], "lr": o.position_lr_init}, {"params": [self.distance], "lr": o.position_lr_init},
            {"params": [self.f_dc], "lr": o.feature_lr}, {"params": [self.f_rest], "lr": o.feature_lr / 20.0},
            {"params": [self.opacity], "lr": o.opacity_lr}, {"params": [self.scaling], "lr": o.scaling_lr},
            {"params": [self.rotation], "lr": o.rotation_lr}], lr=0.0, eps=1e-15, foreach=False, fused=False)
        P = self.bc.shape[0]
        self.max_radii2D = torch.zeros(P, device=device)
        self.grad_accum = torch.zeros(P, 1, device=device)
        self.denom = torch.zeros(P, 1, device=device)

    def setup(self, cams, bg):
        pass

    def step(self, cam, bg, gt):
        from oracle import train_np
        t, o = self.t, self.o
        bc = torch.softmax(self.bc, dim=1)
        xyz = bc[:, 0:1] * t["vertex1"] + bc[:, 1:2] * t["vertex2"] + bc[:, 2:3] * t["vertex3"] \
            + 4.0 * t["r"] * (torch.sigmoid(self.distance) - 0.5) * t["normal"]
        scales = torch.exp(self.scaling)
        rot = torch.nn.functional.normalize(self.rotation, dim=1)
        opac = torch.sigmoid(self.opacity)
        shs = torch.cat((self.f_dc, self.f_rest), dim=1)
        fr = self.rc.RefFrame(bg, xyz.detach(), opac.detach(), cam.world_view_transform, cam.full_proj_transform,
                              cam.camera_center, math.tan(cam.FoVx * 0.5), math.tan(cam.FoVy * 0.5), self.H, self.W, 3,
                              shs=shs.detach(), scales=scales.detach(), rotations=rot.detach(), sync=False)
        img = fr.color.requires_grad_(True)
        photo, _, _ = train_np.photometric_loss(img, gt, o.lambda_dssim)
        photo.backward()
        rg = fr.backward(img.grad, sync=False)
        ab, ac = t["vertex2"] - t["vertex1"], t["vertex3"] - t["vertex1"]
        radius = torch.sqrt(torch.norm(torch.cross(ab, ac, dim=1), dim=1))
        mr = torch.clamp(scales.max(dim=1).values - o.alpha_mrloss * radius, min=0).sum()
        torch.autograd.backward([xyz, scales, rot, opac, shs, mr],
                                [rg["means3D"], rg["scales"], rg["rotations"], rg["opacity"], rg["sh"], None])
        vis = fr.radii > 0
        self.max_radii2D[vis] = torch.maximum(self.max_radii2D[vis], fr.radii[vis].float())
        self.grad_accum[vis] += torch.norm(rg["means2D"][vis, :2], dim=-1, keepdim=True)
        self.denom[vis] += 1
        self.optimizer.step()
        self.optimizer.zero_grad()
        return photo.detach() + mr.detach()

    def check(self):
        pass


def timed(fn, steps, warmup, barrier, pre=None, post=None):
    """W untimed + K timed calls of fn(i); CUDA-event time in ms (this rank).  pre/post fork and join side streams
    inside the timed region (multi-lane forward rendering).  The cyclic garbage collector is paused for the region
    (as timeit does): a generation-2 pass of a torch process takes tens of milliseconds of host time, during which the
    launch queue drains and the device idles -- observed as a rare 2x outlier of one section."""
    gc.collect()
    gc.disable()
    try:
        return _timed(fn, steps, warmup, barrier, pre, post)
    finally:
        gc.enable()


def _timed(fn, steps, warmup, barrier, pre=None, post=None):
    if pre:
        pre()
    for i in range(warmup):
        fn(i)
    if post:
        post()
    torch.cuda.synchronize()
    barrier()
    torch.cuda.synchronize()
    # one event per step boundary (a marker in the stream, no stall): total = first -> last, plus per-step statistics
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    t0 = time.perf_counter()
    marks[0].record()
    if pre:
        pre()
    for i in range(steps):
        fn(warmup + i)
        if not post:
            marks[i + 1].record()
    if post:
        post()
        marks[steps].record()
    timed.last_enqueue_ms = (time.perf_counter() - t0) * 1e3      # host time to enqueue K steps (diagnostic)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    barrier()
    if not post:
        per = sorted(marks[i].elapsed_time(marks[i + 1]) for i in range(steps))
        q = lambda f: per[min(len(per) - 1, int(f * len(per)))]
        timed.last_step_stats = {"median_ms": statistics.median(per), "p10_ms": q(0.10), "p90_ms": q(0.90), "max_ms": per[-1]}
    else:
        timed.last_step_stats = None        # frames in flight on side streams: no per-step boundary on this stream
    return marks[0].elapsed_time(marks[steps]), wall


class OursArm:
    name = "ours"

    def __init__(self, device, scene, W, H):
        from gaussianmesh_b200.renderer import TrainStep, ViewBatchRenderer
        self.device = device
        self.ts = TrainStep(device, scene["means3D"], scene["opacities"], scene["shs"], scene["scales"], scene["rotations"], W, H)
        self.lanes = FORWARD_LANES
        self.vb = ViewBatchRenderer(device, scene["means3D"], scene["opacities"], shs=scene["shs"], scales=scene["scales"],
                                    rotations=scene["rotations"], lanes=self.lanes)
        self.fwd_out = torch.empty(self.lanes, 3, H, W, dtype=torch.float32, device=device)

    def train(self, cam, bg, target):
        return self.ts.step(cam, bg, target)

    def forward(self, cam, bg):
        lane = self.vb._count % self.lanes
        self.vb.render_into(cam, bg, self.fwd_out[lane])
        return self.fwd_out[lane]

    def setup(self, cams, bg):
        # buffer sizing, not warm-up: the binning chunks are allocated once for the largest view of the shard
        self.ts.reserve_for(cams, bg)
        self.vb.reserve_for(cams, bg)

    def check(self):
        bad = self.ts.arena.verify() + self.vb.verify()
        if bad:
            raise RuntimeError(f"arena overflow inside the timed region (frames {bad}); numbers invalid")

    def counters(self):
        return self.ts.arena.last_info


class ReferenceArm:
    """The unmodified reference CUDA rasterizer driven as rasterize_points.py drives it: fresh zero-filled chunks
    per frame, forward_0, host read of num_rendered, forward_1; backward with nine zero-filled gradient tensors.
    L1 loss and its gradient are torch elementwise ops (the reference uses Jittor elementwise ops)."""
    name = "reference"

    def __init__(self, device, scene, W, H):
        import refcuda
        if not refcuda.available():
            raise FileNotFoundError(refcuda.REF_LIB)
        self.rc = refcuda
        self.device, self.scene, self.W, self.H = device, scene, W, H
        self.last_R = None

    def _frame(self, cam, bg):
        s = self.scene
        fr = self.rc.RefFrame(bg, s["means3D"], s["opacities"], cam.world_view_transform, cam.full_proj_transform,
                              cam.camera_center, math.tan(cam.FoVx * 0.5), math.tan(cam.FoVy * 0.5), self.H, self.W, 3,
                              shs=s["shs"], scales=s["scales"], rotations=s["rotations"], sync=False)
        self.last_R = fr.R
        return fr

    def train(self, cam, bg, target):
        fr = self._frame(cam, bg)
        diff = fr.color - target
        loss = diff.abs().mean()
        dL = torch.sign(diff) / diff.numel()
        self.grads = fr.backward(dL, sync=False)
        return loss

    def forward(self, cam, bg):
        return self._frame(cam, bg).color

    def setup(self, cams, bg):
        pass

    def check(self):
        pass

    def counters(self):
        return (self.last_R, None, 0, None)

    # ---- variant (i) of BASELINE.md 2.1: single-call Rasterizer::forward over persistent pre-sized chunks
    def best_setup(self, cams, bg):
        worst = 0
        for cam in cams:
            worst = max(worst, self._frame(cam, bg).R)
        torch.cuda.synchronize()
        s = self.scene
        self.arena = self.rc.RefArena(self.device, s["means3D"].shape[0], self.H, self.W, s["shs"].shape[1], int(worst * 1.05) + 1024)

    def best_train(self, cam, bg, target):
        s = self.scene
        color = self.arena.forward(bg, s["means3D"], s["opacities"], cam.world_view_transform, cam.full_proj_transform,
                                   cam.camera_center, math.tan(cam.FoVx * 0.5), math.tan(cam.FoVy * 0.5), 3, shs=s["shs"],
                                   scales=s["scales"], rotations=s["rotations"])
        diff = color - target
        loss = diff.abs().mean()
        dL = torch.sign(diff) / diff.numel()
        self.arena.backward(dL)
        return loss


_JSON_FD = None


def emit_json(obj) -> None:
    """The ONE JSON line goes to the process's original stdout; everything else any library prints (NCCL's version banner
    is a printf on stdout) was redirected to stderr at start-up."""
    line = (json.dumps(obj) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, line)


def main():
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)                      # C-level and Python-level stdout -> stderr from here on
    args = parse_args()
    if not torch.cuda.is_available():
        emit_json({"error": "no CUDA device: the hot path has no CPU fallback"})
        sys.exit(1)
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    from gaussianmesh_b200.view_shard import ShardContext
    ctx = ShardContext("nccl", device)   # NCCL is used ONLY for the timing barrier and the max over ranks
    rank, world = ctx.rank, ctx.world
    barrier, max_over_ranks = ctx.barrier, ctx.max_over_ranks

    P = args.gaussians
    scene, cams_host, cams, targets_host, targets, cams_packed_host = build_workload(device, P, rank, world)
    bg = torch.zeros(3, dtype=torch.float32, device=device)
    try:
        arm = OursArm(device, scene, WIDTH, HEIGHT) if args.impl == "ours" else ReferenceArm(device, scene, WIDTH, HEIGHT)
    except FileNotFoundError as ex:          # oracle/_ref was not built (it is built by __graft_entry__.build() where the reference lives)
        if rank == 0:
            emit_json({"impl": "reference", "unavailable": f"{ex} is missing: run `make -C oracle ref` where /root/reference exists"})
        ctx.close()
        return
    K, Wm = args.steps, args.warmup
    nv = len(cams)
    # started before the sizing pass so that nvidia-smi's start-up (NVML initialisation) is over when the first timed
    # region begins; only samples taken under load enter the reported median
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if not args.no_presize:
        arm.setup(cams, bg)

    # ---------------------------------------------------------------- (1) device-resident training frames
    # ours reads the resident 8-bit image directly (value / 255 inside the loss kernel); the reference's glue takes the
    # float image its loader produced
    train_targets = build_workload.targets_u8 if args.impl == "ours" else targets

    def train_resident(i):
        arm.train(cams[i % nv], bg, train_targets[i % TARGET_POOL])

    prof = None
    ms_dev, _ = timed(train_resident, K, Wm, barrier)            # the headline region: no stage events
    enqueue_ms = timed.last_enqueue_ms
    step_stats = timed.last_step_stats
    ms_prof = None
    if args.impl == "ours":
        # the same K steps again with every stage bracketed by CUDA events (per-kernel times for the roofline)
        from gaussianmesh_b200 import _lib
        _lib.profile_begin()
        ms_prof, _ = timed(train_resident, K, 0, barrier)
        prof = _lib.profile_end()
    arm.check()
    info = arm.counters()
    ms_dev = max_over_ranks(ms_dev)
    graph_stats = None
    if args.impl == "ours":
        # the same resident step as ONE cudaGraphLaunch per frame (plus the two small device copies that feed the graph's
        # static camera / target buffers): the host-side cost of a step drops to a single launch
        arm.ts.capture(cams[0], bg, train_targets[0])

        def train_graph(i):
            arm.ts.step_graph(cams[i % nv], bg, train_targets[i % TARGET_POOL])
        ms_graph, _ = timed(train_graph, K, Wm, barrier)
        graph_enqueue = timed.last_enqueue_ms
        if arm.ts.verify():
            raise RuntimeError("arena overflow inside the timed CUDA-graph region")
        ms_graph = max_over_ranks(ms_graph)
        graph_stats = {"ms_per_step": ms_graph / K, "value": world * K / (ms_graph * 1e-3), "unit": UNIT,
                       "host_enqueue_ms_per_step": graph_enqueue / K, "step_ms": timed.last_step_stats}

    # ---------------------------------------------------------------- (2) end to end with host inputs
    from gaussianmesh_b200.cameras import DeviceCamera
    from gaussianmesh_b200.feed import HostFrameFeed
    loss_host = torch.zeros(1, dtype=torch.float32).pin_memory()
    if args.impl == "ours":
        # frame i+1's camera and target upload on a copy stream while frame i renders (every copy is still
        # inside the timed region; the loop is primed with frame 0's upload)
        feed = HostFrameFeed(device, [(35,), (3, HEIGHT, WIDTH)], dtypes=[torch.float32, torch.uint8])

        def upload(i):
            feed.push(cams_packed_host[i % nv], targets_host[i % TARGET_POOL])

        def train_e2e(i):
            if feed._pending == 0:
                upload(i)
            cam_dev, target_dev = feed.pop()
            upload(i + 1)
            loss = arm.train(DeviceCamera.from_packed(cams_host[0], cam_dev), bg, target_dev)
            feed.release()
            loss_host.copy_(loss.reshape(1), non_blocking=True)
    else:
        # the reference's glue uploads on the compute stream
        cam_dev = torch.empty(35, dtype=torch.float32, device=device)
        target_dev = torch.empty(3, HEIGHT, WIDTH, dtype=torch.uint8, device=device)
        e2e_cam = DeviceCamera.from_packed(cams_host[0], cam_dev)

        def train_e2e(i):
            cam_dev.copy_(cams_packed_host[i % nv], non_blocking=True)
            target_dev.copy_(targets_host[i % TARGET_POOL], non_blocking=True)
            loss = arm.train(e2e_cam, bg, target_dev.float() / 255.0)       # PILtoTorch: uint8 -> float / 255.0 (torch ops)
            loss_host.copy_(loss.reshape(1), non_blocking=True)

    ms_e2e, wall_e2e = timed(train_e2e, K, Wm, barrier)
    e2e_stats = timed.last_step_stats
    arm.check()
    ms_e2e = max_over_ranks(max(ms_e2e, wall_e2e))
    loss_value = float(loss_host[0])

    # ---------------------------------------------------------------- (3) forward-only view shard
    def fwd(i):
        arm.forward(cams[i % nv], bg)

    vb = getattr(arm, "vb", None)
    ms_fwd, _ = timed(fwd, K, Wm, barrier, pre=vb.begin_batch if vb else None, post=vb.end_batch if vb else None)
    arm.check()
    ms_fwd = max_over_ranks(ms_fwd)

    # ---------------------------------------------------------------- (3b) N-GPU == 1-GPU: every rank renders global view 0
    # (BASELINE.md 3 gate 3): the image bits must not depend on the rank / shard that produced them
    from gaussianmesh_b200 import synthetic as _syn
    from gaussianmesh_b200.cameras import DeviceCamera as _DC
    view0 = _DC.upload(_syn.orbit_cameras(NUM_VIEWS, WIDTH, HEIGHT)[0], device)
    if vb:
        vb.begin_batch()
    img0 = arm.forward(view0, bg)
    if vb:
        vb.end_batch()
    torch.cuda.synchronize()
    bits = img0.contiguous().view(torch.int32).to(torch.int64)
    checksum = (int(bits.sum().item()), int((bits * torch.arange(1, bits.numel() + 1, device=device).view_as(bits) % 1000003).sum().item()))
    shard_checksums = ctx.gather_objects(checksum)
    shard_identical = all(c == shard_checksums[0] for c in shard_checksums)
    arm.check()

    # ---------------------------------------------------------------- (3c) reference only: its best case (variant i)
    ref_best = None
    if args.impl == "reference":
        arm.best_setup(cams, bg)
        ms_best, _ = timed(lambda i: arm.best_train(cams[i % nv], bg, targets[i % TARGET_POOL]), K, Wm, barrier)
        ms_best = max_over_ranks(ms_best)
        ref_best = {"value": world * K / (ms_best * 1e-3), "unit": UNIT, "ms_per_step": ms_best / K,
                    "step_ms": timed.last_step_stats,
                    "what": "BASELINE.md 2.1 variant (i): Rasterizer::forward (single call, its own mid-frame host read) over "
                            "persistent pre-sized chunks, no per-frame allocation or chunk memset; backward into one "
                            "pre-allocated gradient slab zeroed once per frame; torch elementwise L1"}
        del arm.arena

    # ---------------------------------------------------------------- (4) edit path (config 5): deform once, orbit render
    white = torch.ones(3, dtype=torch.float32, device=device)
    if args.impl == "ours":
        obj, (Vd, Rv, Sv), edit_cams, acap_ms = build_edit_workload(device, rank, world)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        ev0.record()
        obj.deform(Vd, Rv, Sv)
        ev1.record()
        torch.cuda.synchronize()
        deform_ms = ev0.elapsed_time(ev1)
        ne = len(edit_cams)
        from gaussianmesh_b200.arena import RenderArena
        edit_arena = RenderArena(device, strict=False)

        def edit_frame(i):
            obj.render_gaussian(edit_cams[i % ne], white, arena=edit_arena)
        edit_frame(0)
        edit_arena.reserve(int(edit_arena.high_water * 1.5))
    else:
        import refcuda
        obj = RefEditObject(device, rank, world)          # torch ops only: the reference arm never maps our library
        edit_cams, deform_ms, acap_ms = obj.cams, obj.deform_ms, None
        ne = len(edit_cams)

        def edit_frame(i):
            cam = edit_cams[i % ne]
            colors = refcuda.edit_colors_torch(obj.deform_pos, cam.camera_center, obj.deform_rot, obj.shs, 3)
            refcuda.RefFrame(white, obj.deform_pos, obj.opacity, cam.world_view_transform, cam.full_proj_transform,
                             cam.camera_center, math.tan(cam.FoVx * 0.5), math.tan(cam.FoVy * 0.5), HEIGHT, WIDTH, 3,
                             colors=colors, cov3D=obj.deform_cov6, M=16, sync=False)
    ms_edit, _ = timed(edit_frame, K, Wm, barrier)
    if args.impl == "ours" and edit_arena.verify():
        raise RuntimeError("edit arena overflow inside the timed region")
    ms_edit = max_over_ranks(ms_edit)
    del obj, edit_cams
    if args.impl == "ours":
        del edit_arena

    # ---------------------------------------------------------------- (5) full training iteration
    it_arrays = build_iteration_workload(P)
    it_arm = (OursIteration if args.impl == "ours" else ReferenceIteration)(device, it_arrays, WIDTH, HEIGHT)
    if not args.no_presize:
        it_arm.setup(cams, bg)

    def train_iteration(i):
        it_arm.step(cams[i % nv], bg, targets[i % TARGET_POOL])

    it_prof = None
    ms_it, _ = timed(train_iteration, K, Wm, barrier)
    it_enqueue_ms = timed.last_enqueue_ms
    if args.impl == "ours":
        from gaussianmesh_b200 import _lib
        _lib.profile_begin()
        ms_it_prof, _ = timed(train_iteration, K, 0, barrier)
        it_prof = _lib.profile_end()
    it_arm.check()
    it_info = it_arm.counters() if args.impl == "ours" else None
    ms_it = max_over_ranks(ms_it)

    # ---------------------------------------------------------------- (6) view-parallel training (N > 1 only)
    view_parallel = None
    if args.impl == "ours" and world > 1:
        from gaussianmesh_b200.renderer import MeshGaussianModel
        from gaussianmesh_b200.training import OptimizationParams
        from gaussianmesh_b200.view_parallel import ViewParallelTrainer
        del it_arm
        view_parallel = {"global_batch_views": world,
                         "workload": "the section-5 iteration with one view per rank per step: gradients averaged over the "
                                     "ranks, one optimizer step per global batch; the densification statistics accumulate per "
                                     "rank and are merged over the ranks when densify_and_prune consumes them "
                                     "(ViewParallelTrainer.merge_stats), not with two collectives every step"}
        for mode in ("nccl", "p2p", "mc"):
            try:
                vp_model = MeshGaussianModel(it_arrays, device, requires_grad=False)
                vp = ViewParallelTrainer(vp_model, OptimizationParams(), WIDTH, HEIGHT, mode=mode)
                if not args.no_presize:
                    vp.reserve_for(cams, bg)
                torch.cuda.synchronize()
                barrier()
                nv0 = nvlink_bytes(local_rank) if rank == 0 else None
                ms_vp, _ = timed(lambda i: vp.step(cams[i % nv], bg, targets[i % TARGET_POOL]), K, Wm, barrier)
                nv1 = nvlink_bytes(local_rank) if rank == 0 else None
                if vp.it.arena.verify():
                    raise RuntimeError("arena overflow inside the timed view-parallel region")
                ms_vp = max_over_ranks(ms_vp)
                # every replica must hold the same parameters after the same global steps
                flat = torch.cat([p_.detach().reshape(-1) for p_ in vp_model.parameters()]).contiguous().view(torch.int32).to(torch.int64)
                psum = (int(flat.sum().item()), int((flat[::97] * 31 % 1000003).sum().item()))
                sums = ctx.gather_objects(psum)
                view_parallel[mode] = {"ms_per_step": ms_vp / K, "views_per_s": world * K / (ms_vp * 1e-3),
                                       "step_ms": timed.last_step_stats,
                                       "replicas_identical": all(c == sums[0] for c in sums)}
                if nv0 is not None and nv1 is not None:
                    # hardware NVLink data counters of rank 0's GPU around the W + K steps of this mode
                    view_parallel[mode]["nvlink_tx_mb_per_step"] = (nv1[0] - nv0[0]) / (K + Wm) / 1e6
                    view_parallel[mode]["nvlink_rx_mb_per_step"] = (nv1[1] - nv0[1]) / (K + Wm) / 1e6
                del vp, vp_model, flat
            except Exception as ex:       # keep the headline numbers if symmetric memory is unavailable on a box
                view_parallel[mode] = {"error": repr(ex)[:300]}
    clocks = sampler.stop() if sampler is not None else None
    replicas_identical = None
    if view_parallel is not None:
        flags = [view_parallel[m].get("replicas_identical") for m in ("nccl", "p2p", "mc") if isinstance(view_parallel.get(m), dict)]
        flags = [f for f in flags if f is not None]
        replicas_identical = all(flags) if flags else None

    if rank != 0:
        ctx.close()
        return

    N = world
    h2d = cams_packed_host[0].numel() * 4 + targets_host[0].numel() * targets_host[0].element_size()
    out = {
        "metric": METRIC, "value": N * K / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": N, "steps": K, "warmup": Wm,
        "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{P} Gaussians (SH degree 3, scale/rotation path), {WIDTH}x{HEIGHT}, forward + L1 to a random "
                               f"8-bit target image + full backward per frame; {NUM_VIEWS}-view orbit sharded over {N} GPU(s) in contiguous "
                               "blocks, no collective on the render path",
                   "gaussians": P, "width": WIDTH, "height": HEIGHT, "views": NUM_VIEWS, "views_per_rank": nv,
                   "l2": "inputs larger than L2 (scene 236 MB + 300 MB gradients + 150 MB binning per frame; a different "
                         "view every step)"},
        "e2e": {"value": N * K / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / K, "loss": loss_value, "step_ms": e2e_stats},
        "step_ms": step_stats,
        "cuda_graph": graph_stats,
        "shard_identical": shard_identical,
        "replicas_identical": replicas_identical,
        "forward": {"value": N * K / (ms_fwd * 1e-3), "unit": UNIT, "ms_per_frame": ms_fwd / K,
                    "frames_in_flight": FORWARD_LANES if args.impl == "ours" else 1},
        "edit": {"value": N * K / (ms_edit * 1e-3), "unit": UNIT, "ms_per_frame": ms_edit / K, "deform_ms": deform_ms,
                 "acap_get_rs_ms": acap_ms,
                 "workload": f"{EDIT_P} mesh-bound Gaussians (5,120-face proxy mesh), deformed once, {EDIT_VIEWS}-frame orbit "
                             f"at {WIDTH}x{HEIGHT}: rotated-direction SH colours + forward with precomputed colour/covariance"},
        "train_iteration": {"value": N * K / (ms_it * 1e-3), "unit": "iterations/s", "ms_per_iteration": ms_it / K,
                            "workload": f"{P} mesh-bound Gaussians (5,120-face proxy mesh) at {WIDTH}x{HEIGHT}: learning-rate "
                                        "schedule, bind + activations, render, (1-l) L1 + l (1-SSIM) + mesh-restrict loss, "
                                        "full backward, densification statistics, Adam over all parameters "
                                        "(train_mesh_gaussian.py:73-147 without densify_and_prune)"},
        "clocks": clocks,
    }
    if view_parallel is not None:
        view_parallel["exchange"] = {"nccl": "ncclAllReduce of the flat gradient vector + replicated one-launch Adam",
                                     "p2p": "gm_adam_step_sharded_p2p: reduce-scatter + Adam + all-gather in one kernel over "
                                            "NVLink peer memory (symmetric memory, P2P loads / stores), optimizer state sharded",
                                     "mc": "gm_adam_step_sharded_mc: the same kernel over the NVSwitch multicast mapping: "
                                           "multimem.ld_reduce (sum inside the switch) + multimem.st (store to all replicas)"}
        out["view_parallel"] = view_parallel
    # ---- cross-arm correctness guard: both arms run the same K + W steps on the same views / targets, so the loss of the
    # last end-to-end step must agree.  Each arm leaves its value in a scratch file; the arm that runs second compares.
    key = f"P{P}_N{N}_K{K}_W{Wm}"
    guard_path = os.path.join(tempfile.gettempdir(), "gm_bench_e2e_loss.json")
    try:
        guard = json.load(open(guard_path)) if os.path.exists(guard_path) else {}
    except (OSError, ValueError):
        guard = {}
    other = guard.get(key, {}).get("ours" if args.impl == "reference" else "reference")
    if other is not None:
        out["e2e"]["other_arm_loss"] = other
        out["e2e"]["loss_matches_other_arm"] = bool(abs(other - loss_value) <= 1e-5)
    guard.setdefault(key, {})[args.impl] = loss_value
    try:
        json.dump(guard, open(guard_path, "w"))
    except OSError:
        pass
    if other is not None and not out["e2e"]["loss_matches_other_arm"]:
        print(f"bench.py: e2e loss {loss_value!r} of arm {args.impl!r} differs from the other arm's {other!r}", file=sys.stderr)

    if args.impl == "reference":
        out["impl"] = "reference"
        out["reference_best"] = ref_best
        out["gpu_launches"] = 0
        out["instances_per_frame"] = info[0]
        out["cpu_baseline"] = {"value": out["value"], "unit": UNIT, "cores": 0, "kind": "reference",
                               "sample": "the reference's own implementation of this path is CUDA: its unmodified "
                                         "cuda_rasterizer sources rebuilt for sm_100a (oracle/_ref), run on the same GPU "
                                         "with the call protocol of its rasterize_points.py, all K steps"}
        out["e2e"] = {k: v for k, v in out["e2e"].items()}      # value, unit, copies, loss, step statistics, cross-arm guard
        emit_json(out)
        ctx.close()
        return

    # ---------------------------------------------------------------- roofline of the dominant kernel
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    need, visible, overflow, cap = info
    tiles = ((WIDTH + 15) // 16) * ((HEIGHT + 15) // 16)
    npx = WIDTH * HEIGHT
    Rn, Vn = float(need), float(visible)
    # algorithmic bytes per launch (DESIGN.md "Algorithmic bytes"), with R = instances this pipeline blends
    alg = {
        "preprocess": 52.0 * P + 259.0 * Vn,
        "emit": 36.0 * Vn + 8.0 * Rn,
        "sort_pack": 16.0 * Rn + 40.0 * Rn,
        "blend_forward": 40.0 * Rn + 20.0 * npx + 8.0 * tiles,
        "blend_backward": 76.0 * Rn + 20.0 * npx + 8.0 * tiles,
        "geometry_backward": 559.0 * Vn + 4.0 * P,
        "l1_loss": 36.0 * npx,
    }
    stages = {k: {"ms_per_launch": v[0] / v[1], "launches": v[1], "share": v[0] / ms_prof} for k, v in prof.items()}
    top = max(stages, key=lambda k: stages[k]["ms_per_launch"] * stages[k]["launches"])
    for k, st in stages.items():
        if k in alg:
            st["algorithmic_bytes"] = alg[k]
            st["achieved_gbs"] = alg[k] / (st["ms_per_launch"] * 1e-3) / 1e9
            st["frac_of_hbm_peak"] = st["achieved_gbs"] / peak
    traffic = None
    tr_path = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.exists(tr_path):
        traffic = json.load(open(tr_path)).get(top)
    out["roofline"] = {"bound": "hbm", "kernel": top, "achieved": stages[top].get("achieved_gbs"), "peak": peak,
                       "peak_source": peak_src, "unit": "GB/s", "frac": stages[top].get("frac_of_hbm_peak"),
                       "traffic": traffic, "algorithmic_bytes": alg.get(top)}
    out["stages"] = stages
    numel_params = float(sum(v.size for k, v in it_arrays.items()
                             if k in ("bc_logits", "distance", "shs", "opacity_logit", "log_scales", "rot_raw")))
    it_alg = {"photometric_loss": 36.0 * npx, "adam": 28.0 * numel_params, "mesh_bind_forward": (64.0 + 32.0 + 44.0) * P,
              "mesh_bind_backward": (64.0 + 32.0 + 44.0 + 44.0) * P, "mesh_restrict_loss": 60.0 * P, "densify_stats": 28.0 * P}
    it_stages = {k: {"ms_per_launch": v[0] / v[1], "launches": v[1], "share": v[0] / ms_it_prof} for k, v in it_prof.items()}
    for k, st in it_stages.items():
        if k in it_alg:
            st["algorithmic_bytes"] = it_alg[k]
            st["achieved_gbs"] = it_alg[k] / (st["ms_per_launch"] * 1e-3) / 1e9
            st["frac_of_hbm_peak"] = st["achieved_gbs"] / peak
    out["train_iteration"]["stages"] = it_stages
    out["train_iteration"]["host_enqueue_ms_per_iteration"] = it_enqueue_ms / K
    out["train_iteration"]["instances_per_frame"] = it_info[0]
    out["train_iteration"]["visible_gaussians"] = it_info[1]
    # depth_histogram + bucket_lut; preprocess + large_tiles (count); emit + large_tiles (placement); bucket_sort_pack + the two
    # size classes of big_bucket_sort_pack; the rest launch one kernel
    kernels_per_stage = {"depth_buckets": 2, "preprocess": 2, "emit": 2, "sort_pack": 3}
    out["gpu_launches"] = int(sum(v[1] * kernels_per_stage.get(k, 1) for k, v in prof.items()))
    out["host_enqueue_ms_per_step"] = enqueue_ms / K
    out["ms_per_step_with_stage_events"] = ms_prof / K
    out["instances_per_frame"] = need
    out["visible_gaussians"] = visible

    # ---------------------------------------------------------------- CPU baseline (bounded sample)
    if N == 1 and not args.no_cpu_baseline:
        try:
            from oracle import cpu_oracle
            out["cpu_baseline"] = cpu_oracle.timed_sample(P, WIDTH, HEIGHT, frames=args.cpu_sample_frames)
        except Exception as ex:   # the baseline is a reported number, never a dependency of the product
            out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {ex!r}"}
    emit_json(out)
    ctx.close()


if __name__ == "__main__":
    main()
