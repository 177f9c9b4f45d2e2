import sys, os, time, pickle
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import numpy as np
from fea_diffusion_b200.workload import build_workload
from fea_diffusion_b200.solver import PackedConditions, pack
t=time.perf_counter(); items,_=build_workload(25,4,64); print("gen %.1fs"%(time.perf_counter()-t))
meshes, samples, index = [], [], {}
for it in items:
    if it.plate not in index:
        index[it.plate] = len(meshes); meshes.append((it.setup.coors, it.setup.conn))
    samples.append((index[it.plate], it.kwargs))
for _ in range(3):
    t=time.perf_counter(); pc=PackedConditions(meshes,samples); dt=time.perf_counter()-t
    print("PackedConditions %d samples: %.2f ms, h2d %.1f MB"%(len(samples),dt*1e3,pc.h2d_bytes/1e6))
t=time.perf_counter(); p=pack([it.setup.sample for it in items]); print("pack(arrays) %.2f ms, h2d %.1f MB"%((time.perf_counter()-t)*1e3,p.h2d_bytes/1e6))
import cProfile, pstats
cProfile.run("PackedConditions(meshes,samples)", "/root/repo/build/pack.prof")
pstats.Stats("/root/repo/build/pack.prof").sort_stats("cumtime").print_stats(14)
