import sys, os, json, numpy as np
sys.path.insert(0, "/root/repo")
from fea_diffusion_b200 import Context, pack
from fea_diffusion_b200.workload import build_workload
items, _ = build_workload(100, 4, 64)
packed = pack([it.setup.sample for it in items])
ctx = Context(0)
with ctx.create_batch(packed) as b:
    b.assemble()
    for _ in range(3):
        b.solve(1e-10, 400000)
    st = b.stats(); r = b.download()
print(json.dumps({"lib": os.environ.get("FEA_B200_LIB", "default"), "solve_ms": st["solve_ms"], "cluster_ms": st["cluster_ms"], "cluster_systems": st["cluster_systems"], "clusters": st["cluster_count"], "size": st["cluster_size"],
                  "conv": int((r.status == 0).sum()), "iters": int(r.iters.sum())}), flush=True)
