#!/usr/bin/env python
"""Offline, CPU: scipy SuperLU solutions of BASELINE config 3 (cantilever L4, gusset L3), stored at
the COARSE mesh's vertices (which keep their indices under uniform refinement) as
tests/golden/c3_lu.npz.  bench.py and tests/test_gpu_large.py compare the GPU's ~1 M-DOF solves with
them.  Takes minutes and several GB of RAM; run once."""
import os, sys, time
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fea_diffusion_b200.workload import large_case
from oracle.fea_oracle import OracleProblem
out = {}
for name, lv in (("cantilever", 4), ("gusset", 3)):
    setup, n0 = large_case(name, lv)
    s = setup.sample
    orc = OracleProblem(setup.coors, setup.conn, num_steps=2)
    orc.fixed_vertex[:] = s.fixed.astype(bool)
    orc.load[:] = s.rhs
    t = time.perf_counter()
    K = orc.stiffness(); b = orc.rhs_final()
    lu = spla.splu(sp.csc_matrix(K))
    x = lu.solve(b)
    r = b - K @ x
    x = x + lu.solve(r)            # one step of iterative refinement
    r = b - K @ x
    print(name, lv, "n=%d" % K.shape[0], "%.0f s" % (time.perf_counter() - t), "relres %.2e" % (np.linalg.norm(r) / np.linalg.norm(b)), flush=True)
    u = np.zeros(2 * len(s.coors)); act = np.repeat(~s.fixed.astype(bool), 2); u[act] = x
    out["%s_L%d_u_coarse" % (name, lv)] = u.reshape(-1, 2)[:n0].copy()
    out["%s_L%d_n_dofs" % (name, lv)] = np.int64(K.shape[0])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "c3_lu.npz"), **out)
