#!/usr/bin/env python
"""Throughput of K device-resident steps for different stream counts / priorities."""
import os, sys, json, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fea_diffusion_b200 import pack
from fea_diffusion_b200.pipeline import Pipeline
from fea_diffusion_b200.workload import build_workload
plates = int(sys.argv[1]) if len(sys.argv) > 1 else 100
K = int(sys.argv[2]) if len(sys.argv) > 2 else 6
items, _ = build_workload(plates, 4, 64)
samples = [it.setup.sample for it in items]
size = max(it.size for it in items); affine = np.stack([it.affine for it in items])
packed = pack(samples)
for S, prio in [(1, False), (2, False), (2, True), (3, False), (3, True), (4, True)]:
    pl = Pipeline(0, S, prio)
    def step(ctx, b):
        b.assemble().solve(1e-10, 20000).rasterize(size, affine, 0.1)
    for rep in range(2):
        batches = [pl.ctxs[j % S].create_batch(packed) for j in range(K)]
        pl.synchronize()
        pl.ctxs[0].event_record(0); t0 = time.perf_counter()
        pl.run(batches, lambda ctx, b: step(ctx, b))
        pl.synchronize(); t1 = time.perf_counter()
        pl.ctxs[0].event_record(1)
        ms = pl.ctxs[0].event_elapsed_ms(0, 1)
        for b in batches: b.destroy()
    print(json.dumps({"streams": S, "prio": prio, "ms_per_step": ms / K, "wall_ms_per_step": 1e3 * (t1 - t0) / K,
                      "solves_per_s": len(samples) * K / (ms * 1e-3)}), flush=True)
    pl.close()
