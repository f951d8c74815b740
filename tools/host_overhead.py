#!/usr/bin/env python
"""Where does the wall time of one device-resident evaluation go?  Wall clock per API call (stream
synchronised after each, so GPU work is included) next to the library's own stage timers."""
import os
import sys
import time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from open_pcc_metric_b200 import _native as N, synth  # noqa: E402

A, B = synth.synth_pair(10, 1_000_000, synth.BASE_SEED + 2, step=2, dedup=False, oversample=4)
dev = torch.device("cuda:0")
dA = [torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (A.points, A.colors, A.normals)]
dB = [torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (B.points, B.colors, B.normals)]
ctx = N.Context(0)
YUV = np.array([[0.25, 0.5, 0.25], [1, 0, -1], [-0.5, 1, -0.5]])
acc = {}


def t(name, fn, sync=True):
    t0 = time.perf_counter()
    r = fn()
    t1 = time.perf_counter()
    if sync:
        ctx.synchronize()
    t2 = time.perf_counter()
    a = acc.setdefault(name, [0.0, 0.0])
    a[0] += t1 - t0
    a[1] += t2 - t0
    return r


for it in range(60):
    if it == 10:
        acc.clear()
    a = t("cloud A", lambda: ctx.cloud(*dA))
    b = t("cloud B", lambda: ctx.cloud(*dB))
    t("build_pair", lambda: ctx.build_pair(a, b))
    t("pair_eval", lambda: ctx.pair_eval(a, b, N.EVAL_D2 | N.EVAL_COLOR, YUV))
    t("close", lambda: (a.close(), b.close()))
n = 50
print(f"{'call':12s} {'host return us':>15s} {'incl. GPU us':>13s}")
for k, (h, g) in acc.items():
    print(f"{k:12s} {h / n * 1e6:15.1f} {g / n * 1e6:13.1f}")
# back-to-back without intermediate synchronisation (what bench.py times)
ctx.synchronize()
t0 = time.perf_counter()
for it in range(50):
    a = ctx.cloud(*dA); b = ctx.cloud(*dB); ctx.build_pair(a, b)
    ctx.pair_eval(a, b, N.EVAL_D2 | N.EVAL_COLOR, YUV); a.close(); b.close()
ctx.synchronize()
print(f"back to back: {(time.perf_counter() - t0) / 50 * 1e6:.1f} us per evaluation")
# host time of each call when nothing but pair_eval waits for the GPU (the enqueue cost the GPU has to be fed with)
acc.clear()
for it in range(60):
    if it == 10:
        acc.clear()
    a = t("cloud A", lambda: ctx.cloud(*dA), sync=False)
    b = t("cloud B", lambda: ctx.cloud(*dB), sync=False)
    t("build_pair", lambda: ctx.build_pair(a, b), sync=False)
    t("pair_eval", lambda: ctx.pair_eval(a, b, N.EVAL_D2 | N.EVAL_COLOR, YUV), sync=False)
    t("close", lambda: (a.close(), b.close()), sync=False)
print("enqueue only (pair_eval includes its one wait):")
for k, (h, g) in acc.items():
    print(f"{k:12s} {h / n * 1e6:15.1f}")
