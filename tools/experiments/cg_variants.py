"""CPU check of the one-hand-over CG variant (beta from |r|^2 - 2 a r.q + a^2 q.q) against the textbook
recurrence on the ill-conditioned bench plate (seed 6) and an ordinary one."""
import sys, os
import numpy as np, scipy.sparse as sp
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from fea_diffusion_b200.workload import plate_conditions
from oracle.fea_oracle import OracleProblem

def system(seed):
    it = plate_conditions(seed, 1, 64)[0][0]
    p = OracleProblem(it.setup.coors, it.setup.conn, **it.kwargs)
    K = p.stiffness().tocsr(); b = p.rhs_final()
    s = 1 / np.sqrt(K.diagonal()); S = sp.diags(s)
    return (S @ K @ S).tocsr(), s * b

def cg_textbook(A, b, rtol=1e-10, maxit=6000):
    x = np.zeros_like(b); r = b.copy(); p = r.copy(); rr = r @ r; r0 = rr
    for k in range(1, maxit + 1):
        q = A @ p; a = rr / (p @ q); x += a * p; r -= a * q
        rr2 = r @ r
        if rr2 <= rtol * rtol * r0: return k, np.linalg.norm(b - A @ x) / np.sqrt(r0)
        p = r + (rr2 / rr) * p; rr = rr2
    return maxit, np.linalg.norm(b - A @ x) / np.sqrt(r0)

def cg_fused(A, b, rtol=1e-10, maxit=6000, variant="expand"):
    x = np.zeros_like(b); r = b.copy(); p = r.copy(); r0 = r @ r
    worst = 0.0
    for k in range(maxit):
        q = A @ p
        pq, qq, rq, rr = p @ q, q @ q, r @ q, r @ r
        if rr <= rtol * rtol * r0: return k, np.linalg.norm(b - A @ x) / np.sqrt(r0), worst
        a = rr / pq
        if variant == "expand": est = rr - 2 * a * rq + a * a * qq
        else: est = a * a * qq - rr          # Saad: uses r.q == p.q
        x += a * p; r -= a * q
        true = r @ r
        worst = max(worst, abs(est - true) / true)
        beta = est / rr if est > 0 else 0.0
        p = r + beta * p
    return maxit, np.linalg.norm(b - A @ x) / np.sqrt(r0), worst

for seed in (6, 3):
    A, b = system(seed)
    print("plate", seed, "n", A.shape[0])
    print("  textbook:", cg_textbook(A, b))
    print("  fused (expansion):", cg_fused(A, b))
    print("  fused (Saad):", cg_fused(A, b, variant="saad"))
