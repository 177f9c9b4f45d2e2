"""Experiment (CPU): why does bench sample 24 stagnate, and does extended-precision iterative
refinement around the fp64 CG fix it?"""
import sys, os, time
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from fea_diffusion_b200.workload import plate_conditions
from oracle.fea_oracle import OracleProblem

def cg(A, b, rtol, maxit, x0=None):
    x = np.zeros_like(b) if x0 is None else x0.copy()
    r = b - A @ x if x0 is not None else b.copy()
    p = r.copy(); rz = r @ r; r0 = rz
    hist = []
    for k in range(1, maxit + 1):
        q = A @ p; a = rz / (p @ q); x += a * p; r -= a * q
        rz2 = r @ r
        if k % 256 == 0: hist.append((k, np.sqrt(rz2 / r0), np.linalg.norm(b - A @ x) / np.sqrt(r0)))
        if rz2 <= rtol * rtol * r0: return x, k, hist
        p = r + (rz2 / rz) * p; rz = rz2
    return x, maxit, hist

plate, cond = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (6, 0)
items, _ = plate_conditions(plate, 4, 64)
it = items[cond]
p = OracleProblem(it.setup.coors, it.setup.conn, **it.kwargs)
K = p.stiffness().tocsr(); b = p.rhs_final()
s = 1 / np.sqrt(K.diagonal()); S = sp.diags(s)
A = (S @ K @ S).tocsr(); bh = s * b
print("n", A.shape[0], "fixed", int(p.fixed_vertex.sum()), "regions", p.n_regions)
lu = spla.splu(sp.csc_matrix(K))
x_lu = lu.solve(b)
# high-precision reference: iterative refinement with longdouble residuals
xr = x_lu.astype(np.longdouble)
Kl = K.astype(np.float64)
for _ in range(5):
    r = (b.astype(np.longdouble) - (sp.csr_matrix(Kl).astype(np.longdouble) @ xr)) if False else None
    # scipy sparse has no longdouble matvec: do it by COO accumulation
    coo = K.tocoo()
    Ax = np.zeros(len(b), np.longdouble); np.add.at(Ax, coo.row, coo.data.astype(np.longdouble) * xr[coo.col])
    r = b.astype(np.longdouble) - Ax
    xr = xr + lu.solve(np.asarray(r, np.float64)).astype(np.longdouble)
x_ref = np.asarray(xr, np.float64)
rel = lambda x: np.linalg.norm(x - x_ref) / np.linalg.norm(x_ref)
print("SuperLU vs refined reference: %.2e" % rel(x_lu))
xh, k, hist = cg(A, bh, 1e-10, 20000)
print("CG iters", k, "true relres %.2e" % (np.linalg.norm(bh - A @ xh) / np.linalg.norm(bh)), "err vs ref %.2e" % rel(s * xh), "vs LU %.2e" % (np.linalg.norm(s*xh - x_lu)/np.linalg.norm(x_lu)))
for h in hist[-6:]: print("   it %d recursive %.2e true %.2e" % h)
print("|x|*|A|*eps/|b| estimate: %.2e" % (np.linalg.norm(xh) * 2.0 * 2.2e-16 / np.linalg.norm(bh)))
# iterative refinement: true residual in extended precision (longdouble here; double-double on the GPU), inner CG fp64
coo = A.tocoo(); dl = coo.data.astype(np.longdouble)
def resid_ld(xhi, xlo):
    Ax = np.zeros(len(bh), np.longdouble)
    np.add.at(Ax, coo.row, dl * (xhi[coo.col].astype(np.longdouble) + xlo[coo.col].astype(np.longdouble)))
    return bh.astype(np.longdouble) - Ax
xhi, xlo = xh.copy(), np.zeros_like(xh)
tot = k
for rnd in range(4):
    r = resid_ld(xhi, xlo); rn = float(np.sqrt(r @ r)) / np.linalg.norm(bh)
    xsum = (xhi.astype(np.longdouble) + xlo)
    print(" round %d: extended-precision true relres %.3e, err vs ref %.2e (total iters %d)" % (rnd, rn, rel(s * np.asarray(xsum, np.float64)), tot))
    if rn <= 1e-10: break
    r64 = np.asarray(r, np.float64)
    inner_tol = max(1e-10 / rn * 0.5, 1e-8)
    d, kk, _ = cg(A, r64, inner_tol, 20000)
    tot += kk
    t = xsum + d.astype(np.longdouble)
    xhi = np.asarray(t, np.float64); xlo = np.asarray(t - xhi.astype(np.longdouble), np.float64)
