#!/usr/bin/env python
"""Dump per-system features and measured PCG iteration counts of bench workloads (for the longest-job-first
estimate of the cluster queues): python tools/experiments/iter_features.py out.npz seed0 [seed0 ...]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fea_diffusion_b200 import Context, pack
from fea_diffusion_b200.workload import build_workload
out = sys.argv[1]
rows = []
ctx = Context(0)
for seed0 in [int(v) for v in sys.argv[2:]]:
    items, _ = build_workload(100, 4, 64, seed0=seed0, workers=os.cpu_count() or 1)
    packed = pack([it.setup.sample for it in items])
    with ctx.create_batch(packed) as b:
        b.assemble().solve(1e-10, 20000)
        r = b.download()
    for i, it in enumerate(items):
        s = it.setup.sample
        xy = np.asarray(s.coors); fx = np.asarray(s.fixed, bool)
        fxy = xy[fx]
        # distance of every free vertex to the nearest fixed one
        d = np.sqrt(((xy[:, None, :] - fxy[None, ::max(1, len(fxy) // 64), :]) ** 2).sum(-1)).min(1)
        h = np.sqrt(np.ptp(xy[:, 0]) * np.ptp(xy[:, 1]) / len(xy))
        D = np.asarray(s.D).reshape(-1, 3, 3)
        rows.append((seed0, i, len(xy), int(fx.sum()), d.max() / h, d.mean() / h, len(D), D[:, 0, 0].max() / D[:, 0, 0].min(),
                     float(np.abs(np.asarray(s.rhs)).sum() > 0), int(r.iters[i]), int(r.status[i])))
np.save(out, np.array(rows, dtype=np.float64))
print("saved", len(rows))
