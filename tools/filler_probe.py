#!/usr/bin/env python
"""Can other kernels use the SMs that a solve leaves idle (12 % of its SM-time)?  A stream of small
low-priority torch kernels ("filler", 148 CTAs x ~20 us each) runs beside the solve of the bench
batch; the filler throughput alone, beside the solve, and the solve time with/without filler tell
whether the hardware places other work while cluster CTAs are pending."""
import os, sys, json, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fea_diffusion_b200 import Context, pack
from fea_diffusion_b200.workload import build_workload
items, _ = build_workload(int(sys.argv[1]) if len(sys.argv) > 1 else 100, 4, 64)
ctx = Context(0)
packed = pack([it.setup.sample for it in items], alloc=ctx.pinned_empty)
lo, hi = torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else (0, -5)
fs = torch.cuda.Stream(priority=0)
x = torch.ones(148 * 1024 * 8, device="cuda")
g = torch.cuda.CUDAGraph()
with torch.cuda.stream(fs):
    for _ in range(3):
        x.mul_(1.0000001)
    fs.synchronize()
    with torch.cuda.graph(g, stream=fs):
        for _ in range(2000):
            x.mul_(1.0000001)
K = 20
def filler():
    with torch.cuda.stream(fs):
        for _ in range(K):
            g.replay()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
filler(); torch.cuda.synchronize()
e0.record(fs); filler(); e1.record(fs); torch.cuda.synchronize()
alone_ms = e0.elapsed_time(e1)
import threading
with ctx.create_batch(packed) as b:
    b.assemble()
    for rep in range(2):
        ctx.event_record(0); b.solve(1e-10, 20000); ctx.event_record(1); ctx.synchronize()
    solo = ctx.event_elapsed_ms(0, 1)
    out = []
    for rep in range(3):
        done = []
        def run():
            time.sleep(0.03)
            ctx.event_record(0); b.solve(1e-10, 20000); ctx.event_record(1); ctx.synchronize(); done.append(ctx.event_elapsed_ms(0, 1))
        t = threading.Thread(target=run)
        t.start()
        e0.record(fs); filler(); e1.record(fs)
        t.join(); torch.cuda.synchronize()
        out.append((round(e0.elapsed_time(e1), 2), round(done[0], 2)))
    print(json.dumps({"filler_ms_alone": alone_ms, "solve_ms_alone": solo, "(filler_ms, solve_ms) together": out,
                      "note": "filler = 40 000 x 148-CTA-wide elementwise kernels on a priority-0 stream"}))
