#!/usr/bin/env python
"""GPU-time breakdown (CUDA events, one synchronisation per stage) of one dataset step on the bench workload:
create_from_conditions -> assemble -> solve -> rasterize -> stage_outputs (region rasters + classifier) -> fetch."""
import os, sys, json, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fea_diffusion_b200 import Context, PackedConditions
from fea_diffusion_b200.workload import build_workload, conditions_of
plates = int(sys.argv[1]) if len(sys.argv) > 1 else 100
items, _ = build_workload(plates, 4, 64, workers=int(os.environ.get("WORKERS", os.cpu_count() or 1)))
meshes, samples = conditions_of(items)
size = max(it.size for it in items); affine = np.stack([it.affine for it in items])
mask = np.array([it.condition == 0 for it in items], np.uint8)
ctx = Context(0)
pc = PackedConditions(meshes, samples, alloc=ctx.pinned_empty)
for rep in range(3):
    ev = lambda i: (ctx.event_record(i), ctx.synchronize())
    ts = [time.perf_counter()]
    gpu = []
    def mark(i):   # (8 event slots per context: two are used alternately)
        ctx.event_record(i & 1); ctx.synchronize(); ts.append(time.perf_counter())
        gpu.append(ctx.event_elapsed_ms((i - 1) & 1, i & 1))
    ctx.event_record(0)
    b = ctx.create_batch_from_conditions(pc); mark(1)
    b.assemble(); mark(2)
    b.solve(1e-10, 20000); mark(3)
    b.rasterize(size, affine, 0.1); mark(4)
    b.stage_outputs(mask); mark(5)
    b.fetch_outputs(); mark(6)
    b.classify(); mark(7)
    if rep == 0:
        n_img = int(pc.n_regions.sum() + mask.sum())
        reg_out = ctx.pinned_empty((n_img, size, size), np.uint8)
    b.rasterize_regions(mask, out=reg_out); mark(8)
    b.destroy()
    names = ("create_from_conditions", "assemble", "solve", "rasterize", "stage_outputs", "fetch_outputs",
             "classify alone", "region rasters alone + D2H into pinned memory")
    print(json.dumps({"gpu_ms": {n: round(gpu[i], 3) for i, n in enumerate(names)},
                      "wall_ms": {n: round(1e3 * (ts[i + 1] - ts[i]), 3) for i, n in enumerate(names)}}), flush=True)
