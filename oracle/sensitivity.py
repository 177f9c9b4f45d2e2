"""How well is a direct solve's answer defined at all?  (TEST INFRASTRUCTURE.)

The parity bar -- relative L2 <= 1e-8 against sfepy's direct solve -- presumes that the direct
solve's own answer is reproducible to better than 1e-8.  For a plate with a weakly held part
(kappa ~ 1e7 and beyond) it is not: the forward error of ANY backward-stable fp64 solve is about
kappa * eps, and merely summing the element contributions in another order (sfepy's C assembly vs
numpy vs the GPU) perturbs K by an ulp.  ``direct_solve_sensitivity`` measures exactly that on the
oracle's own system: by how much scipy's SuperLU solution (what ``ScipyDirect({})`` calls,
reference fea_analysis.py:371-375) moves when every stored entry of K is perturbed by at most half
an ulp, symmetrically.  Parity tests use it as the tolerance band of samples the solver flags as
ill-conditioned (fea_batch_get_refine_rounds > 0); ordinary samples must meet 1e-8 as is.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def direct_solve_sensitivity(K: sp.spmatrix, b: np.ndarray, trials: int = 3, seed: int = 0) -> float:
    K = sp.csr_matrix(K)
    x0 = spla.splu(sp.csc_matrix(K)).solve(b)
    n0 = np.linalg.norm(x0)
    if not n0 > 0:
        return 0.0
    coo = sp.triu(K).tocoo()
    rng = np.random.default_rng(seed)
    worst = 0.0
    for _ in range(trials):
        d = coo.data * (1.0 + 1.1102230246251565e-16 * rng.uniform(-1.0, 1.0, len(coo.data)))
        U = sp.coo_matrix((d, (coo.row, coo.col)), shape=K.shape)
        Kp = (U + sp.triu(U, 1).T).tocsc()
        x = spla.splu(Kp).solve(b)
        worst = max(worst, float(np.linalg.norm(x - x0) / n0))
    return worst
