#!/usr/bin/env python
"""Stage breakdown of one device-resident step on the bench workload + iteration histogram."""
import os, sys, json, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fea_diffusion_b200 import Context, pack
from fea_diffusion_b200.workload import build_workload
plates = int(sys.argv[1]) if len(sys.argv) > 1 else 100
items, _ = build_workload(plates, 4, 64)
samples = [it.setup.sample for it in items]
size = max(it.size for it in items); affine = np.stack([it.affine for it in items])
ctx = Context(0)
packed = pack(samples, alloc=ctx.pinned_empty)
for rep in range(3):
    t0 = time.perf_counter(); ctx.event_record(0)
    b = ctx.create_batch(packed); ctx.event_record(1); ctx.synchronize(); t1 = time.perf_counter()
    b.assemble(); ctx.event_record(2); ctx.synchronize(); t2 = time.perf_counter()
    b.solve(1e-10, 20000); ctx.event_record(3); ctx.synchronize(); t3 = time.perf_counter()
    b.rasterize(size, affine, 0.1); ctx.event_record(4); ctx.synchronize(); t4 = time.perf_counter()
    r = b.download(images=True); t5 = time.perf_counter()
    st = b.stats(); info = b.info()
    b.destroy()
    print(json.dumps({"create_ms": ctx.event_elapsed_ms(0, 1), "assemble_ms": ctx.event_elapsed_ms(1, 2),
                      "solve_ms": ctx.event_elapsed_ms(2, 3), "raster_ms": ctx.event_elapsed_ms(3, 4),
                      "wall_ms": [round(1e3 * (b_ - a_), 2) for a_, b_ in ((t0, t1), (t1, t2), (t2, t3), (t3, t4), (t4, t5))],
                      "stats": st}))
it = np.sort(r.iters)
print("iters percentiles", {p: int(np.percentile(it, p)) for p in (0, 10, 25, 50, 75, 90, 95, 99, 100)}, "mean", it.mean())
print("sum iters", int(it.sum()), "n", len(it), "status", np.bincount(r.status + 1))
nv = np.diff(packed.vtx_off)
print("corr(iters, nv)", np.corrcoef(r.iters, nv)[0, 1])
# cost model: iterations k at which n_active(k) systems are alive
alive = np.array([(it > k).sum() for k in range(0, it.max(), 32)])
print("alive per chunk", alive.tolist())
print("info", info)
