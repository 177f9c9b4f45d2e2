import sys, os, time, cProfile, pstats, threading
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import numpy as np
from fea_diffusion_b200 import Context
from fea_diffusion_b200.workload import build_workload, conditions_of
from fea_diffusion_b200.solver import PackedConditions, PinnedArena
items,_=build_workload(25,4,64)
meshes,samples=conditions_of(items)
samples = samples*4; 
ctx=Context(0)
probe=PackedConditions(meshes,samples)
ar=PinnedArena(ctx, probe.h2d_bytes+(1<<20))
for _ in range(3):
    ar.reset(); t=time.perf_counter(); pc=PackedConditions(meshes,samples,alloc=ar.empty); print("pinned arena: %.2f ms"%((time.perf_counter()-t)*1e3))
for _ in range(2):
    t=time.perf_counter(); pc=PackedConditions(meshes,samples); print("pageable: %.2f ms"%((time.perf_counter()-t)*1e3))
ar.reset(); cProfile.run("PackedConditions(meshes,samples,alloc=ar.empty)","/tmp/pack.prof")
pstats.Stats("/tmp/pack.prof").sort_stats("tottime").print_stats(6)
def work():
    a2=PinnedArena(ctx, probe.h2d_bytes+(1<<20))
    for _ in range(4):
        a2.reset(); t=time.perf_counter(); PackedConditions(meshes,samples,alloc=a2.empty); print("  threaded: %.2f ms"%((time.perf_counter()-t)*1e3))
ts=[threading.Thread(target=work) for _ in range(3)]
[t.start() for t in ts]; [t.join() for t in ts]
