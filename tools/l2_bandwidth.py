#!/usr/bin/env python
"""L2-resident copy and read bandwidth of the GPU (context for the on-chip solver's L2 streaming):
torch copy / sum over buffers that fit the 126 MB L2, CUDA events, best of 20."""
import json
import torch
dev = torch.device("cuda", 0)
out = {}
for mb in (8, 16, 24, 32, 48, 512):
    n = mb * 1024 * 1024 // 4
    a = torch.empty(n, dtype=torch.float32, device=dev).normal_()
    b = torch.empty_like(a)
    best_c = best_r = 1e9
    for _ in range(5):
        b.copy_(a); a.sum()
    for _ in range(20):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(); b.copy_(a); e1.record(); s = a.sum(); e2.record()
        torch.cuda.synchronize()
        best_c = min(best_c, e0.elapsed_time(e1)); best_r = min(best_r, e1.elapsed_time(e2))
    out["%d MB" % mb] = {"copy_GBs_read_plus_write": round(2 * mb / 1024 / (best_c * 1e-3), 1) if best_c else None,
                         "read_GBs": round(mb / 1024 / (best_r * 1e-3), 1)}
print(json.dumps(out))
