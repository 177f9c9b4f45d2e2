"""Does rounding of the Jacobi-scaled matrix (Khat = fl(S K S), unit diagonal) move the solution of the
ill-conditioned bench sample by more than 1e-8?"""
import sys, os
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from fea_diffusion_b200.workload import plate_conditions
from oracle.fea_oracle import OracleProblem
items, _ = plate_conditions(6, 1, 64)
it = items[0]
p = OracleProblem(it.setup.coors, it.setup.conn, **it.kwargs)
K = p.stiffness().tocsr(); b = p.rhs_final()
lu = spla.splu(sp.csc_matrix(K)); x_lu = lu.solve(b)
def refine(A, rhs, lusolve, x, rounds=6):
    coo = A.tocoo(); dl = coo.data.astype(np.longdouble)
    xr = x.astype(np.longdouble)
    for _ in range(rounds):
        Ax = np.zeros(len(rhs), np.longdouble); np.add.at(Ax, coo.row, dl * xr[coo.col])
        r = rhs.astype(np.longdouble) - Ax
        xr = xr + lusolve(np.asarray(r, np.float64)).astype(np.longdouble)
    return np.asarray(xr, np.float64)
x_ref = refine(K, b, lu.solve, x_lu)
rel = lambda a, c: np.linalg.norm(a - c) / np.linalg.norm(c)
print("SuperLU vs exact solution of K x = b: %.2e" % rel(x_lu, x_ref))
d = K.diagonal(); s = 1 / np.sqrt(d)
coo = K.tocoo()
vals = (s[coo.row] * coo.data) * s[coo.col]          # fl(fl(s_i k) s_j), like k_sell_fill
vals[coo.row == coo.col] = 1.0                       # unit diagonal taken as exactly 1
Ah = sp.csr_matrix((vals, (coo.row, coo.col)), shape=K.shape)
bh = s * b
luh = spla.splu(sp.csc_matrix(Ah))
y = refine(Ah, bh, luh.solve, luh.solve(bh))
print("exact solution of the rounded scaled system vs exact of K: %.2e ; vs SuperLU: %.2e" % (rel(s * y, x_ref), rel(s * y, x_lu)))
# perturb K itself by one rounding per entry: the sensitivity of the problem, whoever solves it
rng = np.random.default_rng(0)
for trial in range(3):
    Kp = sp.csr_matrix((coo.data * (1 + 1.1e-16 * rng.uniform(-1, 1, len(coo.data))), (coo.row, coo.col)), shape=K.shape)
    Kp = (Kp + Kp.T) * 0.5
    lup = spla.splu(sp.csc_matrix(Kp)); xp = refine(Kp, b, lup.solve, lup.solve(b))
    print("  K with entries perturbed by <= 1 ulp/2 (symmetric): solution moves %.2e" % rel(xp, x_ref))
