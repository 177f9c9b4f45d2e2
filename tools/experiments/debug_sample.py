import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import numpy as np
from fea_diffusion_b200 import Context, pack
from fea_diffusion_b200.workload import plate_conditions
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 6
it = plate_conditions(seed, 1, 64)[0][0]
ctx = Context(0)
with ctx.create_batch(pack([it.setup.sample])) as b:
    r = b.assemble().solve(1e-10, 50000).download()
    print("status", r.status, "iters", r.iters, "relres", r.relres, "rounds", b.refine_rounds(), b.stats())
ctx.close()
