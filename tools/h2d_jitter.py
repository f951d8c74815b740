#!/usr/bin/env python
"""Is the host->device link steady?  Per-copy times of a 24 MB and a 144 MB pinned H2D copy, 200 repetitions:
min / median / p90 / max.  (bench.py's e2e arm moves 144.6 MB per step; its per-step times on the shared GPU boxes
range from 3.1 ms to tens of ms -- this separates the link from the library.)"""
import time
import torch
for mb in (24, 144):
    h = torch.empty(mb << 20, dtype=torch.uint8, pin_memory=True)
    d = torch.empty(mb << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for i in range(220):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); d.copy_(h, non_blocking=True); e1.record(); torch.cuda.synchronize()
        if i >= 20:
            ts.append(e0.elapsed_time(e1))
    q = sorted(ts)
    print(f"{mb} MB pinned H2D: min {q[0]:.3f}  median {q[len(q)//2]:.3f}  p90 {q[int(len(q)*0.9)]:.3f}  max {q[-1]:.3f} ms  "
          f"(median {mb * 1.048576 / q[len(q)//2]:.1f} GB/s)")
