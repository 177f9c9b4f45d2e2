"""Experiment (CPU, numpy/scipy): iteration counts of Jacobi-PCG vs two-level variants
(aggregation coarse space with rigid-body modes) on bench plates.  Decides whether a coarse-space
correction is worth building into k_pcg_cluster."""
import sys, os, time
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from fea_diffusion_b200.workload import build_workload
from oracle.fea_oracle import OracleProblem, equation_map

def pcg(A, b, apply_M, rtol=1e-10, maxit=20000):
    x = np.zeros_like(b); r = b.copy(); z = apply_M(r); p = z.copy(); rz = r @ z
    r0 = np.sqrt(r @ r)
    for k in range(1, maxit + 1):
        q = A @ p; a = rz / (p @ q); x += a * p; r -= a * q
        if np.sqrt(r @ r) <= rtol * r0: return x, k
        z = apply_M(r); rz2 = r @ z; p = z + (rz2 / rz) * p; rz = rz2
    return x, maxit

def aggregates_tiles(xy, T):
    h = np.sqrt(np.ptp(xy[:, 0]) * np.ptp(xy[:, 1]) / len(xy))
    ix = np.floor((xy[:, 0] - xy[:, 0].min()) / (T * h)).astype(int)
    iy = np.floor((xy[:, 1] - xy[:, 1].min()) / (T * h)).astype(int)
    key = iy * 100000 + ix
    _, agg = np.unique(key, return_inverse=True)
    return agg

def aggregates_rows(xy, G):
    h = np.sqrt(np.ptp(xy[:, 0]) * np.ptp(xy[:, 1]) / len(xy))
    strip = np.floor((xy[:, 1] - xy[:, 1].min()) / h).astype(int)
    order = np.lexsort((xy[:, 0], strip))
    agg = np.empty(len(xy), int); agg[order] = np.arange(len(xy)) // G
    return agg

def coarse_W(xy_act, agg, s, rot=True, scaled=True):
    """W in the Jacobi-scaled space: xhat = S^-1 u.  columns per aggregate: tx, ty, (rot)."""
    n = len(xy_act); na = agg.max() + 1; m = 3 if rot else 2
    rows, cols, vals = [], [], []
    cx = np.bincount(agg, xy_act[:, 0], na) / np.maximum(np.bincount(agg, minlength=na), 1)
    cy = np.bincount(agg, xy_act[:, 1], na) / np.maximum(np.bincount(agg, minlength=na), 1)
    inv = 1.0 / s if scaled else np.ones_like(s)
    v = np.arange(n)
    rows += [2 * v, 2 * v + 1]; cols += [m * agg, m * agg + 1]; vals += [inv[0::2], inv[1::2]]
    if rot:
        dx, dy = xy_act[:, 0] - cx[agg], xy_act[:, 1] - cy[agg]
        rows += [2 * v, 2 * v + 1]; cols += [m * agg + 2, m * agg + 2]; vals += [-dy * inv[0::2], dx * inv[1::2]]
    return sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(2 * n, m * na))

def main():
    nplates = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    items, _ = build_workload(nplates, 4, 64)
    tot = {}
    for it in items:
        p = OracleProblem(it.setup.coors, it.setup.conn, **it.kwargs)
        K = p.stiffness().tocsr(); b = p.rhs_final()
        d = K.diagonal(); s = 1 / np.sqrt(d); S = sp.diags(s)
        A = (S @ K @ S).tocsr(); bh = s * b
        act = ~p.fixed_vertex; xy = p.coors[act]
        res = {}
        _, res["jacobi"] = pcg(A, bh, lambda r: r)
        for name, agg in [("rows128", aggregates_rows(xy, 128)), ("rows256", aggregates_rows(xy, 256)),
                          ("tile8", aggregates_tiles(xy, 8)), ("tile12", aggregates_tiles(xy, 12)), ("tile16", aggregates_tiles(xy, 16)),
                          ("tile24", aggregates_tiles(xy, 24))]:
            for rot in (True, False):
                for scaled in (True,):
                    W = coarse_W(xy, agg, s, rot, scaled)
                    E = (W.T @ A @ W).toarray()
                    # drop null columns (aggregates with singular rigid modes are fine: E SPD if A SPD and W full rank)
                    try:
                        Ei = np.linalg.inv(E)
                    except np.linalg.LinAlgError:
                        continue
                    tag = "%s_%s" % (name, "r" if rot else "t")
                    _, res[tag + "_add"] = pcg(A, bh, lambda r: r + W @ (Ei @ (W.T @ r)))
                    if name in ("tile12", "tile16", "rows256") and rot:
                        # symmetric multiplicative (coarse, smoother, coarse): z = Pc r ; z += (r - A z) ; z += Pc (r - A z)
                        def mult(r, W=W, Ei=Ei):
                            z = W @ (Ei @ (W.T @ r))
                            z = z + (r - A @ z)
                            z = z + W @ (Ei @ (W.T @ (r - A @ z)))
                            return z
                        _, res[tag + "_mult"] = pcg(A, bh, mult)
                    res[tag + "_nc"] = W.shape[1]
        print(it.plate, it.condition, "n=%d" % A.shape[0], " ".join("%s=%d" % kv for kv in sorted(res.items())), flush=True)
        for k, v in res.items(): tot.setdefault(k, []).append(v)
    print("MEAN", " ".join("%s=%.0f" % (k, np.mean(v)) for k, v in sorted(tot.items())))

main()
