"""View-parallel training (SURVEY.md 8e "if training were sharded", 8f-4): one process per GPU, every rank holds the
whole model and renders ITS view of a global batch of `world` views; the first -- and only -- exchange step of this
repository is the gradient average before the optimizer.  The reference has no counterpart (it is single-GPU); the
semantics are those of plain data parallelism over views: the loss is the mean over the batch's views, densification
statistics are accumulated per view as train_mesh_gaussian.py:117-121 does for consecutive iterations.

Three exchange modes:

  "p2p"   gm_adam_step_sharded_p2p: gradient reduce-scatter + Adam + parameter all-gather fused in ONE kernel over
          NVLink peer memory.  Parameters and gradients live in symmetric memory (torch.distributed._symmetric_memory:
          cuMem allocations mapped into every rank); rank r reads shard r of every rank's gradient vector, updates it
          with its shard of the Adam moments (optimizer state is sharded, ZeRO-1 style) and writes the new parameters
          into every rank's parameter vector.  Two device-side barriers bracket the kernel; no host synchronisation.
  "mc"    gm_adam_step_sharded_mc: the same fused kernel over the NVSwitch MULTICAST mapping of the two vectors
          (symmetric memory's multicast_ptr): multimem.ld_reduce sums the N gradient copies inside the switch,
          multimem.st writes the new parameters to all N replicas -- about 2/N of the vector per GPU and direction instead
          of 2 (N-1)/N.  Needs NVLS (an NVSwitch system); `mode="mc"` raises if the mapping is not available.
  "nccl"  the baseline the fused kernel is measured against: ncclAllReduce of the flat gradient vector followed by the
          replicated one-launch Adam (gm_adam_step) on every rank.

`GradientExchange` is the device-agnostic part (torch.distributed only), covered on CPU by a world-size-2 gloo test.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from ._lib import lib, check, AdamSegment, RasterizerError, GM_ERR_BAD_ARGUMENT
from .training import OptimizationParams, TrainingIteration


class GradientExchange:
    """Averages flat gradient vectors and merges per-view densification statistics over the ranks of a process group."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    def average_(self, flat: torch.Tensor) -> torch.Tensor:
        if self.world > 1:
            self.dist.all_reduce(flat, op=self.dist.ReduceOp.SUM, group=self.group)
            flat.mul_(1.0 / self.world)
        return flat

    def merge_stats_(self, max_radii2D: torch.Tensor, grad_accum: torch.Tensor, denom: torch.Tensor,
                     inc_max: torch.Tensor, inc_sum: torch.Tensor) -> None:
        """inc_max [P]: this rank's radii where visible, else 0; inc_sum [2,P]: (|dL/dmean2D.xy|, 1) where visible.
        After the call the persistent statistics hold what `world` consecutive single-view iterations would have
        left (scene/mesh_based_gaussian_model.py:587-589, train_mesh_gaussian.py:117-121)."""
        if self.world > 1:
            self.dist.all_reduce(inc_max, op=self.dist.ReduceOp.MAX, group=self.group)
            self.dist.all_reduce(inc_sum, op=self.dist.ReduceOp.SUM, group=self.group)
        torch.maximum(max_radii2D, inc_max, out=max_radii2D)
        grad_accum.view(-1).add_(inc_sum[0])
        denom.view(-1).add_(inc_sum[1])


class ViewParallelTrainer:
    """One view per rank per step.  `model` must hold identical parameters on every rank."""

    def __init__(self, model, opt: OptimizationParams, W: int, H: int, mode: str = "p2p", group=None,
                 spatial_lr_scale: float = 1.0, merge_stats_every_step: bool = False):
        import torch.distributed as dist
        if mode not in ("auto", "p2p", "mc", "nccl"):
            raise ValueError("mode must be 'auto', 'p2p', 'mc' or 'nccl'")
        auto = mode == "auto"
        if auto:
            mode = "p2p"              # resolved to "mc" below when a multicast mapping exists and the world is large enough
        self.mode = mode
        self.exchange = GradientExchange(group)
        self.world, self.rank = self.exchange.world, self.exchange.rank
        dev = model._bc.device
        self._handles = []
        alloc = None
        fused = mode in ("p2p", "mc")
        if fused and self.world > 1:
            import torch.distributed._symmetric_memory as symm_mem
            self._symm = symm_mem
            self._group = group if group is not None else dist.group.WORLD

            def alloc(n):
                t = symm_mem.empty(n, dtype=torch.float32, device=dev)
                self._handles.append((t, symm_mem.rendezvous(t, self._group)))
                return t
        self.it = TrainingIteration(model, opt, W, H, spatial_lr_scale, alloc=alloc, flat_params=fused)
        it = self.it
        P = it.P
        self._inc_max = torch.zeros(P, dtype=torch.float32, device=dev)
        self._inc_sum = torch.zeros(2, P, dtype=torch.float32, device=dev)
        self.n_step = 0
        # The densification statistics (max radius, gradient accumulator, visit count) are only read when the model is
        # densified, every few hundred steps; max and sum are associative, so each rank accumulates its own views and the
        # ranks are merged when the statistics are consumed (merge_stats(), called by densify_and_prune) instead of with two
        # collectives per step (0.16 ms of a 2.05 ms global step on eight GPUs).
        self.merge_stats_every_step = merge_stats_every_step
        self._stats_pending = False
        self.profile_sections = False     # True: step() brackets its sections with CUDA events (section_times())
        self._marks = []
        if fused:
            if mode == "mc" and self.world == 1:
                raise RasterizerError("ViewParallelTrainer", GM_ERR_BAD_ARGUMENT,
                                      "mode='mc' needs a process group of at least two ranks (multimem addresses)")
            self._bind_flat_buffers()
            lo, hi = self.shard
            n = max(hi - lo, 4)
            self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)          # this rank's shard only
            self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
            if mode == "mc" and (self._grad_mc == 0 or self._param_mc == 0):
                raise RasterizerError("ViewParallelTrainer", GM_ERR_BAD_ARGUMENT,
                                      "no multicast mapping for the symmetric allocations (NVLS unavailable): use mode='p2p'")
            # measured on 8 x B200: multicast 2.16 ms / global step against 2.23 ms for peer loads / stores; on 2 GPUs the
            # peer version wins (1.87 against 2.15 ms): the switch-side reduction pays from four ranks up
            if auto and self.world >= 4 and self._grad_mc != 0 and self._param_mc != 0:
                self.mode = "mc"

    def reserve_for(self, cams: Sequence, bg: torch.Tensor) -> int:
        return self.it.reserve_for(cams, bg)

    # ------------------------------------------------------------------------------------------
    # densification of the replicated / symmetric-memory model (scene/mesh_based_gaussian_model.py:504-585,
    # train_mesh_gaussian.py:119-132).  Every rank holds the same merged statistics, so every rank takes the same
    # decisions; in the fused modes the flat vectors are re-created (a collective rendezvous) and the sharded Adam
    # moments are re-partitioned for the new layout.
    # ------------------------------------------------------------------------------------------
    def _bind_flat_buffers(self) -> None:
        """Shard range, peer / multicast addresses and the barrier handle of the CURRENT flat vectors."""
        it = self.it
        lo, hi = C.c_size_t(), C.c_size_t()
        lib.gm_adam_shard_range(it.flat_numel, self.world, self.rank, C.byref(lo), C.byref(hi))
        self.shard = (lo.value, hi.value)
        if self.world > 1:
            (g_t, g_h), (p_t, p_h) = self._handles[-2], self._handles[-1]             # alloc order: param_grads, flat_parameters
            assert g_t.data_ptr() == it.param_grads.data_ptr() and p_t.data_ptr() == it.flat_parameters.data_ptr()
            self._grad_ptrs = (C.c_void_p * self.world)(*[int(x) for x in g_h.buffer_ptrs])
            self._param_ptrs = (C.c_void_p * self.world)(*[int(x) for x in p_h.buffer_ptrs])
            self._barrier = p_h
            self._grad_mc = int(getattr(g_h, "multicast_ptr", 0) or 0)
            self._param_mc = int(getattr(p_h, "multicast_ptr", 0) or 0)
        else:
            self._grad_ptrs = (C.c_void_p * 1)(it.param_grads.data_ptr())
            self._param_ptrs = (C.c_void_p * 1)(it.flat_parameters.data_ptr())
            self._barrier = None
            self._grad_mc = self._param_mc = 0

    def _gather_moments(self):
        """The full flat Adam moment vectors from the per-rank shards (once per densification).  Shards differ by at most
        one 32-float block, so every rank contributes a buffer of the largest shard's size and the padding is dropped."""
        it = self.it
        ranges = []
        for r in range(self.world):
            lo, hi = C.c_size_t(), C.c_size_t()
            lib.gm_adam_shard_range(it.flat_numel, self.world, r, C.byref(lo), C.byref(hi))
            ranges.append((lo.value, hi.value))
        width = max(hi - lo for lo, hi in ranges)
        full = []
        for mine in (self.exp_avg, self.exp_avg_sq):
            lo, hi = ranges[self.rank]
            send = torch.zeros(width, dtype=torch.float32, device=it.device)
            send[:hi - lo].copy_(mine[:hi - lo])
            recv = torch.empty(self.world * width, dtype=torch.float32, device=it.device)
            if self.world > 1:
                self.exchange.dist.all_gather_into_tensor(recv, send, group=self.exchange.group)
            else:
                recv.copy_(send)
            vec = torch.zeros(it.flat_numel, dtype=torch.float32, device=it.device)
            for r, (a, b) in enumerate(ranges):
                vec[a:b].copy_(recv[r * width:r * width + (b - a)])
            full.append(vec)
        return full

    def densify_and_prune(self, max_grad: float, min_opacity: float = 0.005, extent: float = 0.0, max_screen_size=None,
                          N: int = 4) -> int:
        """TrainingIteration.densify_and_prune for the view-parallel trainer, every mode.  Returns the number of Gaussians
        split (the same on every rank)."""
        it = self.it
        self.merge_stats()
        fused = self.mode in ("p2p", "mc")
        if not fused:
            S = it.densify_and_prune(max_grad, min_opacity, extent, max_screen_size, N)
            if S:
                self._inc_max = torch.zeros(it.P, dtype=torch.float32, device=it.device)
                self._inc_sum = torch.zeros(2, it.P, dtype=torch.float32, device=it.device)
            return S
        grads = it.bc_gradient_accum / it.denom
        grads[grads.isnan()] = 0.0
        sel = grads.reshape(-1) >= max_grad
        S = int(sel.sum().item())
        if S == 0:
            return 0
        keep = ~sel
        old_layout = dict(it.layout)
        old_rows = {gname: getattr(it.model, attr).shape[0] for attr, gname, _ in it.PARAMS}
        full_m, full_v = self._gather_moments()
        n_step = self.n_step
        it.densify_and_split(grads, max_grad, extent, N)      # model surgery + new flat vectors (collective rendezvous)
        self._bind_flat_buffers()
        self._handles = self._handles[-2:]                    # the previous symmetric allocations can go
        # moments of the survivors keep their values, children start at zero (:411-422, cat_tensors_to_optimizer)
        new_m = torch.zeros(it.flat_numel, dtype=torch.float32, device=it.device)
        new_v = torch.zeros_like(new_m)
        kept = int(keep.sum().item())
        for attr, gname, _ in it.PARAMS:
            o_off, o_n = old_layout[gname]
            n_off, _ = it.layout[gname]
            width = o_n // old_rows[gname]
            for src, dst in ((full_m, new_m), (full_v, new_v)):
                dst[n_off:n_off + kept * width].view(kept, width).copy_(src[o_off:o_off + o_n].view(old_rows[gname], width)[keep])
        lo, hi = self.shard
        n = max(hi - lo, 4)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=it.device)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=it.device)
        self.exp_avg[:hi - lo].copy_(new_m[lo:hi])
        self.exp_avg_sq[:hi - lo].copy_(new_v[lo:hi])
        self._inc_max = torch.zeros(it.P, dtype=torch.float32, device=it.device)
        self._inc_sum = torch.zeros(2, it.P, dtype=torch.float32, device=it.device)
        self.n_step = n_step
        return S

    def _device_barrier(self, channel: int) -> None:
        if self._barrier is not None:
            self._barrier.barrier(channel=channel)

    def _mark(self, name: str) -> None:
        """profile_sections: one CUDA event per section boundary of step() (read with section_times())."""
        if self.profile_sections:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self._marks.append((name, ev))

    def section_times(self) -> dict:
        """Mean milliseconds per section of the steps recorded since profile_sections was set (synchronises)."""
        torch.cuda.synchronize(self.it.device)
        acc, cnt = {}, {}
        for (n0, e0), (n1, e1) in zip(self._marks, self._marks[1:]):
            if n1 == "begin":
                continue
            acc[n1] = acc.get(n1, 0.0) + e0.elapsed_time(e1)
            cnt[n1] = cnt.get(n1, 0) + 1
        self._marks = []
        return {k: acc[k] / cnt[k] for k in acc}

    def step(self, cam, bg: torch.Tensor, gt_image: torch.Tensor, iteration: Optional[int] = None) -> torch.Tensor:
        """Enqueue one global step (this rank's view); returns this rank's (photometric loss, L1, SSIM, mrloss)."""
        it = self.it
        self._mark("begin")
        # this rank's statistics go into its own accumulators (zeroed by every merge), merged over the ranks later
        keep = (it.max_radii2D, it.bc_gradient_accum, it.denom)
        it.max_radii2D, it.bc_gradient_accum, it.denom = self._inc_max, self._inc_sum[0], self._inc_sum[1]
        try:
            losses = it.step(cam, bg, gt_image, iteration, optimizer_step=False)
        finally:
            it.max_radii2D, it.bc_gradient_accum, it.denom = keep
        self._mark("render_and_backward")
        self.n_step += 1
        stream = torch.cuda.current_stream(it.device).cuda_stream
        if it.iteration < it.opt.iterations:
            if self.mode == "nccl":
                self.exchange.average_(it.param_grads)
                self._mark("all_reduce")
                it.optimizer.step(it._grad_of)
                self._mark("adam")
            else:
                rows = it.adam_segments()
                segs = (AdamSegment * len(rows))(*[AdamSegment(*r) for r in rows])
                self._device_barrier(0)          # every rank's gradients are complete
                self._mark("barrier_gradients_ready")
                if self.mode == "mc":
                    check(lib.gm_adam_step_sharded_mc(self.world, self.rank, self._grad_mc, self._param_mc,
                                                      it.flat_parameters.data_ptr(), len(rows), segs, it.flat_numel,
                                                      self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), self.n_step, 0.9,
                                                      0.999, 1e-15, stream), "gm_adam_step_sharded_mc")
                else:
                    check(lib.gm_adam_step_sharded_p2p(self.world, self.rank, self._grad_ptrs, self._param_ptrs, len(rows),
                                                       segs, it.flat_numel, self.exp_avg.data_ptr(),
                                                       self.exp_avg_sq.data_ptr(), self.n_step, 0.9, 0.999, 1e-15, stream),
                          "gm_adam_step_sharded_p2p")
                self._mark("exchange_kernel")
                self._device_barrier(1)          # every rank's parameter stores have landed
                self._mark("barrier_parameters_landed")
        if it.iteration < it.opt.densify_until_iter:
            self._stats_pending = True
            if self.merge_stats_every_step:
                self.merge_stats()
                self._mark("merge_stats")
        return losses

    def merge_stats(self) -> None:
        """Fold what every rank accumulated since the last merge into the persistent statistics of `self.it` (collective:
        every rank must call it).  Afterwards they hold what the same views rendered one after the other by a single
        process would have left (scene/mesh_based_gaussian_model.py:587-589, train_mesh_gaussian.py:117-121)."""
        if not self._stats_pending:
            return
        it = self.it
        self.exchange.merge_stats_(it.max_radii2D, it.bc_gradient_accum, it.denom, self._inc_max, self._inc_sum)
        self._inc_max.zero_()
        self._inc_sum.zero_()
        self._stats_pending = False
