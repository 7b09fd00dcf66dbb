"""Copy-engine transport of the tensor-core operands for the sharded path (opt-in: AC_SHARD_PIPELINE=1 and
AC_SHARD_TRANSPORT=symm).

STATUS: written at the end of round 1 after the GPU budget was spent -- NOT yet run on a B200.  It is never
selected by default; DESIGN.md section 8 item 2 is the plan it implements.

Why: NCCL's send/recv kernels occupy SMs while the persistent min-distance GEMM is running.  With the bank
buffers allocated as torch symmetric memory (CUDA VMM allocations mapped into every rank of the node over
NVLink / NVSwitch), a rank PULLS the shards it needs with plain device-to-device copies from the peer mapping:
those run on the copy engines, cost no SM, and each shard gets its own event, so the GEMM launch for shard k
starts as soon as shard k has landed while shards k+1.. are still in flight.

Ordering (all stream-ordered, no host synchronisation):
  barrier A   peers have finished pulling my slice of the PREVIOUS step (their main stream waited for their pull
              events before it reached this barrier)          -> I may overwrite my slice
  local copy  my operands -> my slice of the symmetric buffers
  barrier B   every rank's slice is written                   -> pulls may start
  pulls       side stream, ring order (rank+1, rank+2, ...), one event per source shard
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch


class _EventRequest:
    """Looks like a torch.distributed Work to run_path_sharded: wait() orders the current stream after the pull."""

    def __init__(self, event: torch.cuda.Event):
        self.event = event

    def wait(self) -> None:
        torch.cuda.current_stream().wait_event(self.event)


class SymmetricBank:
    """Persistent symmetric buffers for (hi, lo, n2) of the WHOLE bank, one set per (rows, D, dtype, lo?, group)."""

    _cache = {}

    @classmethod
    def get(cls, total_rows: int, D: int, dtype: torch.dtype, want_lo: bool, device, group) -> "SymmetricBank":
        import torch.distributed as dist

        g = group if group is not None else dist.group.WORLD
        key = (total_rows, D, dtype, want_lo, str(device), g.group_name)
        if key not in cls._cache:
            cls._cache[key] = cls(total_rows, D, dtype, want_lo, device, g)
        return cls._cache[key]

    def __init__(self, total_rows: int, D: int, dtype: torch.dtype, want_lo: bool, device, group):
        import torch.distributed._symmetric_memory as symm

        self.group = group
        self.hi = symm.empty((total_rows, D), dtype=dtype, device=device)
        self.lo = symm.empty((total_rows, D), dtype=dtype, device=device) if want_lo else None
        self.n2 = symm.empty((total_rows,), dtype=torch.float32, device=device)
        self.handles = [None if t is None else symm.rendezvous(t, group) for t in (self.hi, self.lo, self.n2)]   # collective
        self.side = torch.cuda.Stream(device=device)

    def start(self, hi: torch.Tensor, lo: Optional[torch.Tensor], n2: torch.Tensor, bounds: Sequence[Tuple[int, int]], P: int,
              need_rank: Sequence[int], rank: int, world: int):
        """Returns ((hi_buf, lo_buf, n2_buf), [(source rank, [request]), ...]) like distributed.start_shard_pipeline."""
        main = torch.cuda.current_stream()
        h0 = self.handles[0]
        a, b = bounds[rank]
        h0.barrier(channel=0)
        for buf, t in ((self.hi, hi), (self.lo, lo), (self.n2, n2)):
            if buf is not None:
                buf[a * P : b * P].copy_(t)
        h0.barrier(channel=1)
        self.side.wait_stream(main)
        steps: List[Tuple[int, list]] = []
        with torch.cuda.stream(self.side):
            for k in range(1, world):
                src = (rank + k) % world
                if src not in need_rank:
                    continue
                sa, sb = bounds[src]
                for buf, hdl in zip((self.hi, self.lo, self.n2), self.handles):
                    if buf is None:
                        continue
                    peer = hdl.get_buffer(src, tuple(buf.shape), buf.dtype)
                    buf[sa * P : sb * P].copy_(peer[sa * P : sb * P], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.side)
                steps.append((src, [_EventRequest(ev)]))
        return (self.hi, self.lo, self.n2), steps
