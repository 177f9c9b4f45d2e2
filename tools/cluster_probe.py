#!/usr/bin/env python
"""Small batch through the on-chip PCG path (for ncu captures of k_pcg_cluster)."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fea_diffusion_b200 import Context, pack
from fea_diffusion_b200.workload import build_workload
plates = int(sys.argv[1]) if len(sys.argv) > 1 else 4
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
items, _ = build_workload(plates, 4, 64)
samples = [it.setup.sample for it in items]
ctx = Context(0)
packed = pack(samples)
with ctx.create_batch(packed) as b:
    b.assemble()
    for _ in range(reps):
        b.solve(1e-10, 20000)
        st = b.stats()
        print(json.dumps({k: st[k] for k in ("iterations", "n_converged", "solve_ms", "cluster_systems", "cluster_count",
                                            "cluster_iterations", "cluster_ms")}),
              "us/iter/cluster", 1e3 * st["cluster_ms"] * min(st["cluster_count"], st["cluster_systems"]) / max(1, st["cluster_iterations"]))
    r = b.download()
    print("iters", r.iters.tolist()[:16], "nv", np.diff(packed.vtx_off).tolist()[:16])
ctx.close()
