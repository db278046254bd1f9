"""Persistent scratch chunks for sync-free rendering.

The reference allocates the geometry / binning / image chunks afresh every frame and needs a
blocking device->host read of the instance count in the middle of the forward to size the binning
chunk (dgr/rasterize_points.py:118-121,185-194; rasterizer_impl.cu:409-411).  A `RenderArena`
keeps the three chunks alive across frames, sizes the binning chunk from a high-water mark of the
instance counts seen so far, and calls the single-call `gm_forward`, which never synchronises:
each frame's counters land in a ring of pinned host words and are inspected when a later frame is
submitted (or in `verify()`).

Overflow (a view that needs more instances than the chunk holds) is never silent: the frame is
recorded in `overflowed`, `verify()` reports it, and `forward()` raises on the next submission
unless `strict=False` (batch renderers re-render those frames after growing the arena).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from ._lib import lib, check, RasterizerError, GM_ERR_BINNING_OVERFLOW


class RenderArena:
    """Caller-owned scratch for one stream of frames on one device.

    A frame's chunks are overwritten by the next `forward`; run a frame's backward before
    submitting the next frame on the same arena (training does), or use one arena per frame in
    flight.
    """

    RING = 64

    def __init__(self, device, instances: int = 0, headroom: float = 1.25, strict: bool = True):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RasterizerError("RenderArena", -2, "a CUDA device is required (there is no CPU path)")
        self.headroom = float(headroom)
        self.strict = strict
        self.geom: Optional[torch.Tensor] = None
        self.image: Optional[torch.Tensor] = None
        self.binning: Optional[torch.Tensor] = None
        self.capacity = 0                      # instances the binning chunk holds (as the library counts)
        self.high_water = 0
        self._want = int(instances)
        self._info = torch.zeros(self.RING, 4, dtype=torch.int32).pin_memory()
        self._events: List[Optional[torch.cuda.Event]] = [None] * self.RING
        self._frame_of_slot = [-1] * self.RING
        self._next = 0
        self.frames = 0
        self.overflowed: List[int] = []        # frame numbers that did not fit
        self.last_info: Optional[Tuple[int, int, int, int]] = None

    # ------------------------------------------------------------------ sizing
    def _ensure(self, P: int, N: int) -> None:
        g = int(lib.gm_required_geom(P))
        if self.geom is None or self.geom.numel() < g:
            self.geom = torch.empty(g, dtype=torch.uint8, device=self.device)
        i = int(lib.gm_required_image(N))
        if self.image is None or self.image.numel() < i:
            self.image = torch.empty(i, dtype=torch.uint8, device=self.device)

    def reserve(self, instances: int) -> None:
        """Make the binning chunk hold at least `instances` (Gaussian, tile) instances."""
        instances = int(instances)
        if instances <= self.capacity:
            return
        if self.capacity:
            instances = max(instances, int(self.capacity * 1.5))     # geometric growth: few reallocations
        nbytes = int(lib.gm_required_binning(instances))
        self.binning = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        self.capacity = int(lib.gm_binning_capacity(nbytes))

    def reserve_for_views(self, P: int, D: int, M: int, background: torch.Tensor, W: int, H: int, view_args_list) -> int:
        """Size the binning chunk for a known set of views before rendering them: one gm_forward_0 per view
        (preprocess + count, one blocking 4-byte read each), then a single allocation.  Returns the largest count."""
        self._ensure(P, W * H)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        tmp_radii = torch.empty(P, dtype=torch.int32, device=self.device)
        worst = 0
        for view_args in view_args_list:
            n = check(lib.gm_forward_0(self.geom.data_ptr(), P, D, M, background.data_ptr(), W, H, *view_args, 0,
                                       tmp_radii.data_ptr(), 0, stream), "gm_forward_0")
            worst = max(worst, n)
        self.high_water = max(self.high_water, worst)
        self._want = max(self._want, int(worst * self.headroom) + 1024)
        self.reserve(self._want)
        return worst

    # ------------------------------------------------------------------ counters
    def _retire(self, slot: int) -> None:
        ev = self._events[slot]
        if ev is None:
            return
        need, visible, overflow, cap = (int(v) & 0xFFFFFFFF for v in self._info[slot].tolist())
        self.last_info = (need, visible, overflow, cap)
        self.high_water = max(self.high_water, need)
        if overflow:
            self.overflowed.append(self._frame_of_slot[slot])
        self._events[slot] = None

    def poll(self, block: bool = False) -> None:
        """Fold finished frames' counters into the high-water mark (non-blocking unless `block`)."""
        for slot in range(self.RING):
            ev = self._events[slot]
            if ev is None:
                continue
            if block:
                ev.synchronize()
            if ev.query():
                self._retire(slot)

    def verify(self) -> List[int]:
        """Block until every submitted frame is done; return (and clear) the frames that overflowed.

        The arena has already been grown so that re-rendering them fits."""
        self.poll(block=True)
        bad, self.overflowed = self.overflowed, []
        if bad:
            self.reserve(int(self.high_water * self.headroom) + 1024)
        return bad

    # ------------------------------------------------------------------ forward
    def forward(self, P: int, D: int, M: int, background: torch.Tensor, W: int, H: int, view_args: tuple,
                prefiltered: bool, debug: bool, stream: int, out_color: Optional[torch.Tensor] = None,
                radii: Optional[torch.Tensor] = None, epilogue=None):
        """gm_forward over the arena's chunks.  Returns the tuple RasterizeGaussiansCUDA returns, with
        num_rendered = the arena capacity in instances (what gm_backward needs as R).  `epilogue` (a
        _lib.ForwardEpilogue) folds the L1 loss and / or the clearing of the gradient accumulators into the blend kernel."""
        self._ensure(P, W * H)
        self.poll()
        if self.overflowed and self.strict:
            bad = self.verify()
            raise RasterizerError("RenderArena.forward", GM_ERR_BINNING_OVERFLOW,
                                  f"frames {bad} needed more instances than the arena held; the arena has been "
                                  f"grown to {self.capacity} -- re-render them")
        if self.capacity == 0 and self._want == 0:
            # first frame: size the binning chunk with the two-phase protocol once (one blocking read)
            tmp_radii = torch.empty(P, dtype=torch.int32, device=self.device)
            n = check(lib.gm_forward_0(self.geom.data_ptr(), P, D, M, background.data_ptr(), W, H, *view_args,
                                       int(prefiltered), tmp_radii.data_ptr(), int(debug), stream), "gm_forward_0")
            self.high_water = max(self.high_water, n)
            self._want = int(n * self.headroom) + 1024
        # grow when the largest frame seen so far comes within 8 % of the capacity -- not whenever a new high-water
        # mark dents the headroom: a re-allocation is a cudaMalloc of hundreds of MB (milliseconds of host time during
        # which the launch queue drains), and in training every parameter update moves the instance counts a little
        want = self._want
        if self.high_water and self.high_water * 1.08 + 1024 > self.capacity:
            want = max(want, int(self.high_water * self.headroom) + 1024)
        if want > self.capacity:
            self.reserve(want)

        if out_color is None:
            out_color = torch.empty(3, H, W, dtype=torch.float32, device=self.device)
        if radii is None:
            radii = torch.empty(P, dtype=torch.int32, device=self.device)
        slot = self._next
        if self._events[slot] is not None:      # ring wrapped: wait for that old frame
            self._events[slot].synchronize()
            self._retire(slot)
        self._next = (slot + 1) % self.RING
        check(lib.gm_forward_ex(self.geom.data_ptr(), self.binning.data_ptr(), self.binning.numel(),
                                self.image.data_ptr(), P, D, M, background.data_ptr(), W, H, *view_args,
                                int(prefiltered), out_color.data_ptr(), radii.data_ptr(), int(debug),
                                self._info[slot].data_ptr(), epilogue, stream), "gm_forward")
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._events[slot] = ev
        self._frame_of_slot[slot] = self.frames
        self.frames += 1
        return self.capacity, out_color, radii, self.geom, self.binning, self.image
