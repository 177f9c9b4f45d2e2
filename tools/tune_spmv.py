#!/usr/bin/env python
"""Time the PCG kernels for several FEA_SPMV_VARIANT settings on the bench workload
(all systems kept active: rtol = 0, fixed iteration count)."""
import os, sys, json, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fea_diffusion_b200 import Context, pack
from fea_diffusion_b200.workload import build_workload

plates = int(sys.argv[1]) if len(sys.argv) > 1 else 100
variants = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0]
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 640
items, _ = build_workload(plates, 4, 64)
samples = [it.setup.sample for it in items]
for v in variants:
    ctx = Context(0)
    ctx.set_option("pcg_path", 1)       # the streaming kernels (the on-chip path has no SpMV launch)
    ctx.set_option("spmv_variant", v)
    packed = pack(samples)
    with ctx.create_batch(packed) as b:
        b.assemble()
        info = b.info()
        for _ in range(2):
            b.solve(0.0, iters)
        st = b.stats()
    nn, nnz, n = info["n_active_dofs"], info["nnz"], len(samples)
    alg = 12 * nnz + 4 * (nn + n) + 16 * nn
    print(json.dumps({"variant": v, "spmv_ms": st["spmv_ms_avg"], "update_ms": st["update_ms_avg"],
                      "timed": st["spmv_launches_timed"], "solve_ms": st["solve_ms"],
                      "ms_per_iter": st["solve_ms"] / iters,
                      "alg_GBs": alg / st["spmv_ms_avg"] / 1e6, "sell_blocks": info["sell_blocks"]}), flush=True)
    ctx.close()
