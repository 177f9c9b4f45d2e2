#!/usr/bin/env python
"""Per-sample PCG iteration counts of the bench workload with the features a work predictor could
use (tuning of the longest-job-first queues): gpurun_out/iters.npz"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fea_diffusion_b200 import Context, pack
from fea_diffusion_b200.workload import build_workload
plates = int(sys.argv[1]) if len(sys.argv) > 1 else 100
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
items, _ = build_workload(plates, 4, 64, seed0=seed0)
samples = [it.setup.sample for it in items]
ctx = Context(0)
packed = pack(samples)
with ctx.create_batch(packed) as b:
    r = b.assemble().solve(1e-10, 20000).download()
feat = []
for s in samples:
    co = np.asarray(s.coors); D = np.asarray(s.D).reshape(-1, 9)
    fx = np.asarray(s.fixed).astype(bool)
    ext = co.max(0) - co.min(0)
    # spread of the fixed vertices (a clamp along a whole edge pins more than a cluster of points)
    fext = (co[fx].max(0) - co[fx].min(0)) if fx.any() else np.zeros(2)
    cen = co.mean(0); fcen = co[fx].mean(0) if fx.any() else cen
    feat.append([len(co), fx.sum(), len(D), D[:, 8].min(), D[:, 8].max(), ext[0], ext[1], fext[0], fext[1],
                 np.linalg.norm(fcen - cen), np.abs(np.asarray(s.rhs)).sum()])
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez(os.path.join(ROOT, "gpurun_out", "iters_%d.npz" % seed0), iters=r.iters, status=r.status, feat=np.array(feat, float))
print("saved", len(samples), "samples; mean iters", r.iters.mean())
