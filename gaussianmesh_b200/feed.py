"""Double-buffered host -> device feed for per-frame inputs (camera matrices, target image).

The reference uploads each frame's camera / ground-truth tensors on the compute stream, so the copy and the
render serialise.  `HostFrameFeed` stages frame i+1 on a separate copy stream while frame i renders: two device
slots, one "ready" event per slot (copy -> compute) and one "free" event per slot (compute -> copy).  Host
sources must be pinned for the copies to be asynchronous.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch


class HostFrameFeed:
    def __init__(self, device, shapes: Sequence[Tuple[int, ...]], dtype=torch.float32, slots: int = 2, dtypes=None):
        """`dtypes` gives one dtype per shape (default: `dtype` for all), e.g. float32 camera + uint8 target image."""
        self.device = torch.device(device)
        self.copy_stream = torch.cuda.Stream(self.device)
        self.slots = slots
        dtypes = list(dtypes) if dtypes is not None else [dtype] * len(shapes)
        self.buffers: List[List[torch.Tensor]] = [[torch.empty(s, dtype=dt, device=self.device) for s, dt in zip(shapes, dtypes)]
                                                  for _ in range(slots)]
        self.ready = [torch.cuda.Event() for _ in range(slots)]
        self.free = [None] * slots
        self._head = 0      # next slot to fill
        self._tail = 0      # next slot to consume
        self._pending = 0

    def push(self, *host_tensors: torch.Tensor) -> None:
        """Enqueue the upload of one frame's inputs (in the order of `shapes`) on the copy stream."""
        if self._pending >= self.slots:
            raise RuntimeError("HostFrameFeed: all slots in flight; pop() before pushing again")
        slot = self._head
        with torch.cuda.stream(self.copy_stream):
            if self.free[slot] is not None:
                self.copy_stream.wait_event(self.free[slot])       # the render that used this slot has finished
            for dst, src in zip(self.buffers[slot], host_tensors):
                dst.copy_(src, non_blocking=True)
            self.ready[slot].record(self.copy_stream)
        self._head = (slot + 1) % self.slots
        self._pending += 1

    def pop(self) -> List[torch.Tensor]:
        """Device tensors of the oldest pushed frame; the current stream waits for their upload."""
        if self._pending == 0:
            raise RuntimeError("HostFrameFeed: nothing pushed")
        slot = self._tail
        torch.cuda.current_stream(self.device).wait_event(self.ready[slot])
        self._tail = (slot + 1) % self.slots
        self._pending -= 1
        self._last = slot
        return self.buffers[slot]

    def release(self) -> None:
        """Call after enqueueing the work that reads the tensors of the last pop()."""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.free[self._last] = ev
