"""Parity on the bench workload itself (BASELINE config 2, scaled down to 12 plates x 4 conditions):
every sample of a CUDA batch against the CPU oracle's direct solve -- displacements, ranges.txt
values and the two step-1 images -- plus the statistics bench.py reports."""
import numpy as np
import pytest

from fea_diffusion_b200 import Context, pack
from fea_diffusion_b200._capi import SAMPLE_CONVERGED, SAMPLE_STAGNATED
from fea_diffusion_b200.workload import build_workload
from oracle import raster_oracle as ro
from oracle.fea_oracle import OracleProblem

pytestmark = pytest.mark.gpu


def test_every_sample_of_a_workload_batch_matches_the_oracle():
    items, rejected = build_workload(12, 4, 64, seed0=31)
    assert len(items) == 48 and rejected > 0          # the sampler's ill-posed draws are filtered out
    ctx = Context(0)
    try:
        packed = pack([it.setup.sample for it in items])
        size = max(it.size for it in items)
        affine = np.stack([it.affine for it in items])
        t1 = 0.1
        res = ctx.solve_batch(packed, 1e-10, 50000, image_size=size, affine=affine, value_scale=t1)
    finally:
        ctx.close()
    us = packed.split_vertices(res.u)
    assert res.stats["cluster_systems"] == 48         # plate-sized systems: all on the on-chip path
    worst_u, worst_px, n_checked = 0.0, 0, 0
    for i, it in enumerate(items):
        orc = OracleProblem(it.setup.coors, it.setup.conn, num_steps=11, **it.kwargs)
        assert orc.classify()["well_posed"] == 1
        assert res.status[i] in (SAMPLE_CONVERGED, SAMPLE_STAGNATED)
        if res.status[i] != SAMPLE_CONVERGED:          # flagged as too ill-conditioned: not a parity sample
            assert res.relres[i] > 1e-9
            continue
        u = orc.solve("best")
        if not np.any(u[-1]):                           # every loaded vertex is constrained: u == 0 exactly
            assert not np.any(us[i]) and res.iters[i] == 0
            continue
        err = float(np.linalg.norm(us[i] - u[-1]) / np.linalg.norm(u[-1]))
        worst_u = max(worst_u, err)
        assert err <= 1e-8, (i, err)
        for c in range(2):                              # ranges.txt of every step within 1e-8 relative
            for k in (1, 5, 10):
                lo, hi = 0.1 * k * res.ranges[i, 2 * c], 0.1 * k * res.ranges[i, 2 * c + 1]
                span = max(abs(u[k][:, c].min()), abs(u[k][:, c].max()))
                assert abs(lo - u[k][:, c].min()) <= 1e-8 * span and abs(hi - u[k][:, c].max()) <= 1e-8 * span
            ref = ro.rasterize_scalar(orc.coors, orc.conn, u[1][:, c], size, it.affine)
            d = int(np.abs(ref.astype(int) - res.images[i, c].astype(int)).max())
            worst_px = max(worst_px, d)
            assert d <= 1, (i, c, d)                    # images within 1 LSB
        n_checked += 1
    assert n_checked >= 44
    print("workload parity: %d samples, max rel-L2 %.2e, max pixel diff %d" % (n_checked, worst_u, worst_px))
