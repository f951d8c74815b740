#!/usr/bin/env python
"""Wall-clock timeline of ONE end-to-end evaluation from pinned host float64 arrays (what bench.py's e2e arm
times): when does each API call return, and when is the GPU done behind it?  Two passes: calls back to back
(as the product runs), and with a device synchronisation after every call (cost of each piece alone)."""
import os
import sys
import time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from open_pcc_metric_b200 import _native as N, synth  # noqa: E402
from open_pcc_metric_b200.calculator import MetricCalculator  # noqa: E402
from open_pcc_metric_b200.cloud_pair import CloudPair  # noqa: E402
from open_pcc_metric_b200.options import CalculateOptions, transform_options  # noqa: E402

A, B = synth.synth_pair(10, 1_000_000, synth.BASE_SEED + 2, step=2, dedup=False, oversample=4)


def pinned(a):
    t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True)
    t.numpy()[...] = a
    return t.numpy()


hA = synth.Cloud(pinned(A.points), pinned(A.colors), pinned(A.normals))
hB = synth.Cloud(pinned(B.points), pinned(B.colors), pinned(B.normals))
ctx = N.Context(0)
opts = transform_options(CalculateOptions(color="yuv", hausdorff=False, point_to_plane=True))
flush = torch.empty(64 << 20, dtype=torch.float32, device="cuda")


def one(sync_each, log):
    flush.zero_()
    torch.cuda.synchronize()
    t0 = time.perf_counter()

    def mark(name):
        t1 = time.perf_counter()
        if sync_each:
            torch.cuda.synchronize()
        log.setdefault(name, []).append(((t1 - t0) * 1e3, (time.perf_counter() - t0) * 1e3))

    da = ctx.cloud(hA.points); mark("cloud A (coords)")
    db = ctx.cloud(hB.points); mark("cloud B (coords)")
    da.attach(hA.colors, hA.normals); mark("attach A")
    db.attach(hB.colors, hB.normals); mark("attach B")
    ctx.build_pair(da, db); mark("build_pair")
    YUV = np.array([[0.25, 0.5, 0.25], [1, 0, -1], [-0.5, 1, -0.5]])
    ctx.pair_eval(da, db, N.EVAL_D2 | N.EVAL_COLOR, YUV); mark("pair_eval (waits)")
    da.self_nn_minmax(); mark("self_nn A")
    db.self_nn_minmax(); mark("self_nn B")
    da.close(); db.close(); mark("close")
    torch.cuda.synchronize()
    log.setdefault("TOTAL", []).append(((time.perf_counter() - t0) * 1e3,) * 2)


for sync_each in (False, True):
    log = {}
    for it in range(8):
        one(sync_each, log)
    print("== device synchronised after every call" if sync_each else "== back to back")
    for k, v in log.items():
        v = v[3:]
        print(f"{k:22s} returned at {np.mean([x[0] for x in v]):8.3f} ms   done at {np.mean([x[1] for x in v]):8.3f} ms")

# the public surface
for it in range(6):
    flush.zero_()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pair = CloudPair(hA, hB, ctx=ctx, peak="resolution", resolution_bits=10)
    t1 = time.perf_counter()
    out = MetricCalculator(pair).calculate(opts).as_dict()
    t2 = time.perf_counter()
    pair.close()
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    if it >= 3:
        print(f"CloudPair() {1e3 * (t1 - t0):.3f} ms, calculate() {1e3 * (t2 - t1):.3f} ms, close+sync {1e3 * (t3 - t2):.3f} ms, total {1e3 * (t3 - t0):.3f} ms")

# outliers: 300 evaluations back to back; for the slow ones, which call held the host?
names = ["cloud A", "cloud B", "attach A", "attach B", "build_pair", "pair_eval", "self_nn A", "self_nn B", "close"]
YUV = np.array([[0.25, 0.5, 0.25], [1, 0, -1], [-0.5, 1, -0.5]])
rows = []
for it in range(300):
    flush.zero_()
    torch.cuda.synchronize()
    t = [time.perf_counter()]
    da = ctx.cloud(hA.points); t.append(time.perf_counter())
    db = ctx.cloud(hB.points); t.append(time.perf_counter())
    da.attach(hA.colors, hA.normals); t.append(time.perf_counter())
    db.attach(hB.colors, hB.normals); t.append(time.perf_counter())
    ctx.build_pair(da, db); t.append(time.perf_counter())
    ctx.pair_eval(da, db, N.EVAL_D2 | N.EVAL_COLOR, YUV); t.append(time.perf_counter())
    da.self_nn_minmax(); t.append(time.perf_counter())
    db.self_nn_minmax(); t.append(time.perf_counter())
    da.close(); db.close(); t.append(time.perf_counter())
    rows.append([1e3 * (b - a) for a, b in zip(t, t[1:])])
tot = np.array([sum(r) for r in rows])
med = float(np.median(tot))
print(f"300 evaluations: min {tot.min():.3f}  median {med:.3f}  p90 {np.percentile(tot, 90):.3f}  max {tot.max():.3f} ms")
for i in np.argsort(-tot)[:6]:
    if tot[i] > 1.5 * med:
        print(f"  step {i}: {tot[i]:.3f} ms = " + ", ".join(f"{n} {v:.3f}" for n, v in zip(names, rows[i]) if v > 0.05))
