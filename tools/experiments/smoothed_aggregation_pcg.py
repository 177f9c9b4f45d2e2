"""Experiment (CPU, numpy/scipy): two-level PCG with (smoothed) aggregation coarse spaces of rigid-body modes on bench plates,
V(1,1) cycles with damped-Jacobi smoothing and additive variants (1.0-1.65 x fewer SpMV-equivalents than
Jacobi-PCG -- not worth two more hand-overs per iteration on chip)."""
import sys, os, time
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from fea_diffusion_b200.workload import build_workload
from oracle.fea_oracle import OracleProblem

def pcg(A, b, apply_M, rtol=1e-10, maxit=5000):
    x = np.zeros_like(b); r = b.copy(); z = apply_M(r); p = z.copy(); rz = r @ z
    r0 = np.sqrt(r @ r)
    for k in range(1, maxit + 1):
        q = A @ p; a = rz / (p @ q); x += a * p; r -= a * q
        if np.sqrt(r @ r) <= rtol * r0: return x, k
        z = apply_M(r); rz2 = r @ z; p = z + (rz2 / rz) * p; rz = rz2
    return x, maxit

def tiles(xy, T):
    h = np.sqrt(np.ptp(xy[:, 0]) * np.ptp(xy[:, 1]) / len(xy))
    ix = np.floor((xy[:, 0] - xy[:, 0].min()) / (T * h)).astype(int)
    iy = np.floor((xy[:, 1] - xy[:, 1].min()) / (T * h)).astype(int)
    _, agg = np.unique(iy * 100000 + ix, return_inverse=True)
    return agg

def tentative(xy, agg, s):
    n = len(xy); na = agg.max() + 1
    cnt = np.maximum(np.bincount(agg, minlength=na), 1)
    cx = np.bincount(agg, xy[:, 0], na) / cnt; cy = np.bincount(agg, xy[:, 1], na) / cnt
    inv = 1.0 / s
    v = np.arange(n)
    dx, dy = xy[:, 0] - cx[agg], xy[:, 1] - cy[agg]
    rows = [2*v, 2*v+1, 2*v, 2*v+1]; cols = [3*agg, 3*agg+1, 3*agg+2, 3*agg+2]
    vals = [inv[0::2], inv[1::2], -dy*inv[0::2], dx*inv[1::2]]
    W = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(2*n, 3*na))
    # normalise columns
    nrm = np.sqrt(np.asarray(W.multiply(W).sum(0)).ravel()); nrm[nrm == 0] = 1
    return W @ sp.diags(1/nrm)

def main():
    nplates = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    items, _ = build_workload(nplates, 4, 64)
    tot = {}
    for it in items:
        p = OracleProblem(it.setup.coors, it.setup.conn, **it.kwargs)
        K = p.stiffness().tocsr(); b = p.rhs_final()
        d = K.diagonal(); s = 1/np.sqrt(d); S = sp.diags(s)
        A = (S @ K @ S).tocsr(); bh = s * b
        act = ~p.fixed_vertex; xy = p.coors[act]
        rho = spla.eigsh(A, k=1, which='LA', return_eigenvectors=False, tol=1e-3)[0]
        res = {"n": A.shape[0]}
        _, res["jac"] = pcg(A, bh, lambda r: r)
        for T in (6, 8, 12):
            agg = tiles(xy, T)
            W = tentative(xy, agg, s)
            for smooth in (0, 1):
                P = W if not smooth else (W - (4/(3*rho)) * (A @ W))
                E = (P.T @ A @ P).toarray()
                Ei = np.linalg.pinv(E, hermitian=True)
                nnzP = P.nnz / P.shape[0]
                for nu, om in ((1, 2/3*2/rho*1.0),):
                    omega = 1.0 / rho * 4/3
                    def vcyc(r, P=P, Ei=Ei, omega=omega, nu=nu):
                        z = omega * r
                        for _ in range(nu-1): z = z + omega * (r - A @ z)
                        z = z + P @ (Ei @ (P.T @ (r - A @ z)))
                        for _ in range(nu): z = z + omega * (r - A @ z)
                        return z
                    _, k = pcg(A, bh, vcyc)
                    res["T%d_s%d_v%d" % (T, smooth, nu)] = k
                # additive: z = r + P Ei P^T r   (one hand-over)
                _, k = pcg(A, bh, lambda r, P=P, Ei=Ei: r + P @ (Ei @ (P.T @ r)))
                res["T%d_s%d_add" % (T, smooth)] = k
                res["T%d_s%d_nc" % (T, smooth)] = P.shape[1]
                res["T%d_s%d_nzr" % (T, smooth)] = round(nnzP, 1)
        print(it.plate, it.condition, " ".join("%s=%s" % kv for kv in res.items()), flush=True)
        for k, v in res.items(): tot.setdefault(k, []).append(v)
    print("MEAN", " ".join("%s=%.0f" % (k, np.mean(v)) for k, v in tot.items()))
main()
