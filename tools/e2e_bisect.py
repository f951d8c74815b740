#!/usr/bin/env python
"""bench.py's e2e loop shows rare stalls of 30-800 ms that tools/e2e_timeline.py's loop does not: which ingredient?"""
import os, sys, time, gc
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from open_pcc_metric_b200 import _native as N, synth
from open_pcc_metric_b200.calculator import MetricCalculator
from open_pcc_metric_b200.cloud_pair import CloudPair
from open_pcc_metric_b200.options import CalculateOptions, transform_options
A, B = synth.synth_pair(10, 1_000_000, synth.BASE_SEED + 2, step=2, dedup=False, oversample=4)
def pinned(a):
    t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True); t.numpy()[...] = a; return t.numpy()
hA = synth.Cloud(pinned(A.points), pinned(A.colors), pinned(A.normals))
hB = synth.Cloud(pinned(B.points), pinned(B.colors), pinned(B.normals))
opts = CalculateOptions(color="yuv", hausdorff=False, point_to_plane=True)
dev = torch.device("cuda:0")

def run(label, own_stream, sync_between, events, flush_mb, nvml=False, gc_off=False, n=150):
    stream = torch.cuda.Stream(device=dev)
    ctx = N.Context(0) if own_stream else N.Context(0, stream.cuda_stream)
    flush = torch.empty(flush_mb << 20, dtype=torch.uint8, device=dev)
    if gc_off: gc.disable()
    ts = []
    with torch.cuda.stream(stream):
        for it in range(n):
            flush.zero_()
            if sync_between: torch.cuda.synchronize()
            if events:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); e0.record(stream)
            t0 = time.perf_counter()
            pair = CloudPair(hA, hB, ctx=ctx, peak="resolution", resolution_bits=10)
            out = MetricCalculator(pair).calculate(transform_options(opts)).as_dict()
            pair.close()
            if events: e1.record(stream)
            ts.append((time.perf_counter() - t0) * 1e3)
        torch.cuda.synchronize()
    gc.enable()
    q = sorted(ts[10:])
    print(f"{label:60s} min {q[0]:.3f} median {q[len(q)//2]:.3f} p90 {q[int(len(q)*.9)]:.3f} max {q[-1]:.3f} ms", flush=True)
    ctx.close()

run("own stream, sync between, no events, 64MB flush", True, True, False, 64)
run("own stream, NO sync between", True, False, False, 64)
run("torch stream, sync between", False, True, False, 64)
run("torch stream, no sync, events, 256MB flush (= bench)", False, False, True, 256)
run("same, gc disabled", False, False, True, 256, gc_off=True)
run("own stream, no sync, events, 256MB flush", True, False, True, 256)
run("torch stream, no sync, no events, 256MB", False, False, False, 256)
run("torch stream, no sync, events, 64MB", False, False, True, 64)
