#!/usr/bin/env python
"""BASELINE config 3: cantilever / gusset at ~1 M DOFs, single-GPU large-solve path (streaming PCG)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from fea_diffusion_b200 import Context, pack
from fea_diffusion_b200.host import MeshTopology, ProblemSetup
from fea_diffusion_b200.workload import refine_uniform
F = np.load(os.path.join(ROOT, "tests", "golden", "fixtures.npz"))
ctx = Context(0)
for name, levels in (("cantilever", 4), ("gusset", 3)):
    co, cn = refine_uniform(F[name + "_coors"], F[name + "_conn"], levels)
    setup = ProblemSetup(co, cn)
    topo = MeshTopology(setup.conn, len(co))
    if name == "cantilever":
        fixed = topo.facet_vertices(np.flatnonzero(co[:, 0] < 0.01)); loads = [(np.array([3]), (0.0, -1000.0))]
    else:
        fixed = topo.facet_vertices(np.flatnonzero((co[:, 1] < 0.01) | (co[:, 0] < 0.01)))
        loads = [(topo.facet_vertices(np.flatnonzero(co[:, 0] > 0.99)), (1000.0, 0.0))]
    setup.sample.fixed[:] = False; setup.sample.fixed[fixed] = True; setup.sample.rhs[:] = 0
    for v, m in loads: setup.sample.rhs[v] += m
    packed = pack([setup.sample])
    for rep in range(2):
        t0 = time.perf_counter()
        with ctx.create_batch(packed) as b:
            b.assemble(); t1 = time.perf_counter()
            b.solve(1e-10, 400000); ctx.synchronize(); t2 = time.perf_counter()
            st, info, r = b.stats(), b.info(), b.download()
    n, nnz = info["n_active_dofs"], info["nnz"]
    alg = 12 * nnz + 4 * (n + 1) + 16 * n
    print(json.dumps({"mesh": "%s L%d" % (name, levels), "n_dofs": n, "nnz": nnz, "iterations": st["iterations"],
                      "status": int(r.status[0]), "relres": float(r.relres[0]), "assemble_ms": 1e3 * (t1 - t0),
                      "solve_ms": st["solve_ms"], "spmv_ms": st["spmv_ms_avg"], "update_ms": st["update_ms_avg"],
                      "spmv_alg_GBs": alg / st["spmv_ms_avg"] / 1e6 if st["spmv_ms_avg"] else None,
                      "us_per_iteration": 1e3 * st["solve_ms"] / max(1, st["iterations"])}), flush=True)
ctx.close()
