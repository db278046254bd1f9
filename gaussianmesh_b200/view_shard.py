"""One process per GPU, independent views per process, no collective on the render path.

`ShardContext` is the only place torch.distributed is touched: a barrier before/after a timed region and
a MAX reduction of per-rank times (bench.py), plus an optional gather of small per-view records.  The
backend is "nccl" on GPUs and "gloo" in the CPU tests; with world_size == 1 nothing is initialised.
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch


def shard_views(num_views: int, world_size: int, rank: int) -> range:
    """Contiguous blocks, sizes differing by at most one (100 views / 8 ranks -> 13,13,13,13,12,12,12,12;
    SURVEY.md 8e).  Views are independent: no exchange step, no collective."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad world_size / rank")
    if num_views < 0:
        raise ValueError("bad num_views")
    base, extra = divmod(num_views, world_size)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


class ShardContext:
    def __init__(self, backend: Optional[str] = None, device: Optional[torch.device] = None):
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.device = device
        self._dist = None
        if self.world > 1:
            import torch.distributed as dist
            if backend is None:
                backend = "nccl" if (device is not None and device.type == "cuda") else "gloo"
            kw = {"device_id": device} if backend == "nccl" and device is not None else {}
            dist.init_process_group(backend, rank=self.rank, world_size=self.world, **kw)
            self._dist = dist

    def views(self, num_views: int) -> range:
        return shard_views(num_views, self.world, self.rank)

    def barrier(self) -> None:
        if self._dist is not None:
            self._dist.barrier()

    def max_over_ranks(self, x: float) -> float:
        if self._dist is None:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device=self.device if self.device is not None else "cpu")
        self._dist.all_reduce(t, op=self._dist.ReduceOp.MAX)
        return float(t.item())

    def gather_objects(self, obj) -> List:
        """Small per-rank records (view ids, checksums) to every rank; not used on the render path."""
        if self._dist is None:
            return [obj]
        out = [None] * self.world
        self._dist.all_gather_object(out, obj)
        return out

    def close(self) -> None:
        if self._dist is not None:
            self._dist.destroy_process_group()
            self._dist = None
