"""Several contexts (streams) on one GPU working through a queue of batches.

A lock-step batch ends in a tail where only the slowest-converging systems are still iterating and
the GPU is mostly idle; host polls and PCIe copies leave further gaps.  Running two or three
batches at once on separate streams fills those gaps with the bulk phase of another batch.
Streams get distinct priorities so that identical batches do not march in phase.
Worker threads only make ctypes calls (the GIL is released inside them).
"""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor
from typing import Callable, List, Optional, Sequence

from .solver import Context


class Pipeline:
    def __init__(self, device: int = 0, n_streams: int = 2, staggered_priorities: bool = False,
                 first: Optional[Context] = None):
        """``first`` adopts an existing context as stream 0 (closed with the pipeline)."""
        self.device = device
        self.ctxs: List[Context] = [first] if first is not None else []
        while len(self.ctxs) < n_streams:
            i = len(self.ctxs)
            self.ctxs.append(Context(device, priority=(-i if staggered_priorities else 0)))
        self.pool = ThreadPoolExecutor(max_workers=n_streams)

    @property
    def n_streams(self) -> int:
        return len(self.ctxs)

    def synchronize(self):
        for c in self.ctxs:
            c.synchronize()

    def run(self, jobs: Sequence, fn: Callable):
        """Calls fn(ctx, job) for every job; jobs are dealt round-robin to the streams and each
        stream processes its share in order.  Returns results in job order."""
        n = self.n_streams
        out = [None] * len(jobs)

        def work(i):
            ctx = self.ctxs[i]
            for j in range(i, len(jobs), n):
                out[j] = fn(ctx, jobs[j])

        futs = [self.pool.submit(work, i) for i in range(n)]
        for f in futs:
            f.result()
        return out

    def join_into(self, ctx: Context):
        """Stream-order ``ctx`` after everything submitted so far on every stream of the pipeline
        (so one CUDA event on ``ctx`` closes a region timed across all of them)."""
        for c in self.ctxs:
            if c is not ctx:
                ctx.wait_for(c)

    def kernel_launches(self) -> int:
        return sum(c.kernel_launches() for c in self.ctxs)

    def close(self):
        self.pool.shutdown(wait=True)
        for c in self.ctxs:
            c.close()
        self.ctxs = []
