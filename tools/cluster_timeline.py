#!/usr/bin/env python
"""Timeline of one batch's on-chip solves from a FEA_CLUSTER_TRACE build:
    tools/build_variant.sh trace -DFEA_CLUSTER_ACCOUNT -DFEA_CLUSTER_TRACE
    FEA_B200_LIB=build/variants/lib_trace.so FEA_CLUSTER_TRACE_FILE=/tmp/trace.txt python tools/cluster_timeline.py <seed0> [opt=val ...]
Prints, per 1 ms bin, the SMs held by every cluster class, and per class the first start / last end."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fea_diffusion_b200 import Context, pack
from fea_diffusion_b200.workload import build_workload
seed0 = int(sys.argv[1]) if len(sys.argv) > 1 else 0
items, _ = build_workload(100, 4, 64, seed0=seed0, workers=os.cpu_count() or 1)
ctx = Context(0)
for k, v in (a.split("=") for a in sys.argv[2:]):
    ctx.set_option(k, int(v))
packed = pack([it.setup.sample for it in items], alloc=ctx.pinned_empty)
path = os.environ["FEA_CLUSTER_TRACE_FILE"]
with ctx.create_batch(packed) as b:
    b.assemble()
    for _ in range(3):
        b.solve(1e-10, 20000)
    ms = b.stats()["cluster_ms"]
tr = np.loadtxt(path, dtype=np.int64)
t0 = tr[:, 1].min()
st, en, cl, sm, it = (tr[:, 1] - t0) * 1e-6, (tr[:, 2] - t0) * 1e-6, tr[:, 3], tr[:, 4], tr[:, 5]
print("cluster_ms %.2f  systems %d  span %.2f ms" % (ms, len(tr), en.max()))
classes = sorted(set(cl.tolist()))
for c in classes:
    m = cl == c
    print("class %d: %3d systems, first start %.2f, last start %.2f, last end %.2f ms, SM-ms %.0f" %
          (c, m.sum(), st[m].min(), st[m].max(), en[m].max(), float(((en[m] - st[m]) * c).sum())))
nb = int(np.ceil(en.max()))
print("ms   " + " ".join("cl%d" % c for c in classes) + "  SMs busy")
for k in range(nb):
    row = []
    for c in classes:
        m = cl == c
        ov = np.clip(np.minimum(en[m], k + 1) - np.maximum(st[m], k), 0, None).sum() * c
        row.append(ov)
    print("%3d  " % k + " ".join("%3.0f" % r for r in row) + "   %3.0f" % sum(row))
