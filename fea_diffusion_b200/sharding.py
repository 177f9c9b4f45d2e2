"""Multi-GPU plumbing: one process per GPU, plates dealt to ranks, no collective on the data path.

Plate-condition samples are independent (reference ``datagen/generate.py:56,83`` are plain loops), so
the only cross-rank traffic is bookkeeping: a barrier around timed regions, the max of the ranks'
device times, and (for tests/reports) a gather of per-shard digests.  ``torch.distributed`` carries
that -- NCCL on the GPU box, gloo in CPU tests.
"""
from __future__ import annotations

import hashlib
import os
from dataclasses import dataclass
from typing import Iterable, List, Optional, Sequence

import numpy as np


@dataclass
class DistEnv:
    rank: int = 0
    world: int = 1
    local_rank: int = 0

    @staticmethod
    def from_env() -> "DistEnv":
        return DistEnv(int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
                       int(os.environ.get("LOCAL_RANK", "0")))


def plate_shard(num_plates: int, rank: int, world: int, start: int = 0) -> List[int]:
    """Plates of a fixed-size job owned by ``rank``: round-robin by plate index, so that all
    conditions of a plate (which share its mesh) stay on one GPU and the assignment of plate p
    does not depend on how many plates there are."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    return list(range(start + rank, start + num_plates, world))


def weak_scaling_seed(seed: int, rank: int) -> int:
    """First plate seed of ``rank`` when every rank processes its own full-size workload."""
    return seed + 100000 * rank


def reduce_scalar(value: float, op: str = "max", device=None) -> float:
    """max / sum of a host scalar over all ranks (identity without an initialised process group)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
    return float(t.item())


def sample_digest(arrays: Iterable[np.ndarray]) -> str:
    """sha256 over the raw bytes of a sample's arrays (inputs or outputs): equality of digests is
    the byte-identity check between a 1-rank and an N-rank run."""
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode())
        h.update(str(a.shape).encode())
        h.update(a.tobytes())
    return h.hexdigest()


def gather_objects(obj, dst_all: bool = True) -> Sequence:
    """All ranks' ``obj`` in rank order (a one-element list without a process group)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [obj]
    out: List[Optional[object]] = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out
