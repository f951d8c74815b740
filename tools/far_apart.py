#!/usr/bin/env python
"""Worst case for the ring walk: two surfaces that do not overlap at all (offset clouds)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from open_pcc_metric_b200 import _native as N, synth
ctx = N.Context(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
A = synth.synth_vox(10, n, 3, with_colors=False, with_normals=False, oversample=4).points
for off in (0, 50, 500, 5000):
    B = A + np.array([off, 0, off // 2], dtype=np.float64)
    a, b = ctx.cloud(A), ctx.cloud(B)
    ctx.build_pair(a, b)
    ctx.synchronize(); t = time.perf_counter()
    idx, d2 = ctx.nn(a, b)
    dt = time.perf_counter() - t
    print(f"offset {off:5d}: nn of {len(A)} queries in {dt*1e3:8.1f} ms   mean d2 {d2.mean():.1f}", flush=True)
    a.close(); b.close()
