"""N > 1 host path on CPU: world_size-2 gloo process group (127.0.0.1), plates dealt to ranks,
per-plate inputs byte-identical to the single-process job, max/sum reductions used by bench.py."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from fea_diffusion_b200 import pack
from fea_diffusion_b200.sharding import (gather_objects, plate_shard, reduce_scalar, sample_digest,
                                         weak_scaling_seed)
from fea_diffusion_b200.workload import plate_conditions

N_PLATES, CONDS, MESH = 5, 2, 6e-2


def _plate_digest(p):
    items, _ = plate_conditions(p, CONDS, 64, mesh_size=MESH)
    out = []
    for it in items:
        s = it.setup.sample
        out.append(sample_digest([s.coors, s.conn, s.cell_region, s.D, np.asarray(s.fixed, np.uint8), s.rhs, it.affine]))
    return out


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = plate_shard(N_PLATES, rank, world)
        digests = {p: _plate_digest(p) for p in mine}
        # the batch a rank would hand to its GPU: offsets consistent with its own samples only
        samples = [it.setup.sample for p in mine for it in plate_conditions(p, CONDS, 64, mesh_size=MESH)[0]]
        packed = pack(samples)
        assert packed.n == len(mine) * CONDS and packed.vtx_off[-1] == sum(len(s.coors) for s in samples)
        everyone = gather_objects(digests)
        t_max = reduce_scalar(10.0 + rank, "max")
        n_sum = reduce_scalar(float(packed.n), "sum")
        dist.barrier()
        if rank == 0:
            q.put((everyone, t_max, n_sum))
    finally:
        dist.destroy_process_group()


def test_shards_partition_the_job():
    for world in (1, 2, 3, 4, 8):
        owned = [plate_shard(11, r, world) for r in range(world)]
        assert sorted(p for o in owned for p in o) == list(range(11))
        assert max(len(o) for o in owned) - min(len(o) for o in owned) <= 1
    assert plate_shard(10, 1, 4, start=100) == [101, 105, 109]
    with pytest.raises(ValueError):
        plate_shard(4, 2, 2)
    assert weak_scaling_seed(7, 0) == 7 and weak_scaling_seed(7, 3) - weak_scaling_seed(7, 2) == 100000


def test_world_size_2_gloo_matches_single_process():
    serial = {p: _plate_digest(p) for p in range(N_PLATES)}
    assert reduce_scalar(3.5, "max") == 3.5 and gather_objects("x") == ["x"]   # no process group
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    everyone, t_max, n_sum = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert len(everyone) == 2 and set(everyone[0]) == {0, 2, 4} and set(everyone[1]) == {1, 3}
    merged = {**everyone[0], **everyone[1]}
    assert merged == serial            # byte-identical inputs whatever the world size
    assert t_max == 11.0 and n_sum == N_PLATES * CONDS
