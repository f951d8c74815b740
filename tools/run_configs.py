#!/usr/bin/env python
"""Time the other BASELINE.json configurations on one GPU (not the bench metric; used to
find scale problems):  python tools/run_configs.py [c3|c5|c1] [points]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from open_pcc_metric_b200 import _native as N, synth  # noqa: E402
from open_pcc_metric_b200.calculator import MetricCalculator  # noqa: E402
from open_pcc_metric_b200.cloud_pair import CloudPair  # noqa: E402
from open_pcc_metric_b200.options import CalculateOptions, transform_options  # noqa: E402


def timed(ctx, fn):
    ctx.synchronize()
    t = time.perf_counter()
    out = fn()
    ctx.synchronize()
    return out, (time.perf_counter() - t) * 1e3


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "c3"
    ctx = N.Context(0)
    ctx.set_profiling(2)
    if which == "c3":      # vox12 ~4M without normals: kNN+PCA normals then D1/D2/Hausdorff
        n = int(sys.argv[2]) if len(sys.argv) > 2 else 4_000_000
        A = synth.synth_vox(12, n, synth.BASE_SEED + 3, with_colors=False, with_normals=False, oversample=3)
        B = synth.degrade(A, 2, synth.BASE_SEED + 3, 12, dedup=False)
        opts = CalculateOptions(color=None, hausdorff=True, point_to_plane=True)
        peak = dict(peak="resolution", resolution_bits=12)
    elif which == "c5":    # float LiDAR-style
        n = int(sys.argv[2]) if len(sys.argv) > 2 else 5_000_000
        A, B = synth.synth_lidar(n, synth.BASE_SEED + 5)
        m = min(len(A), len(B))
        A, B = synth.Cloud(A.points[:m]), synth.Cloud(B.points[:m])
        opts = CalculateOptions(color=None, hausdorff=True, point_to_plane=True)
        peak = dict(peak="aabb_diag")
    else:                  # c1: vox10 100k D1
        A, B = synth.synth_pair(10, 100_000, synth.BASE_SEED + 1, with_colors=False)
        opts = CalculateOptions()
        peak = dict(peak="resolution", resolution_bits=10)
    print(which, "points", len(A), len(B), flush=True)
    for it in range(2):
        ctx.reset_timings()
        (pair, res), ms = timed(ctx, lambda: (lambda p: (p, MetricCalculator(p).calculate(transform_options(opts)).as_dict()))(
            CloudPair(A, B, ctx=ctx, **peak)))
        tm = ctx.timings()
        print(f"iter {it}: total {ms:.1f} ms  kind {pair.kind}  cell {pair._dev[0].info().cell_size}/{pair._dev[1].info().cell_size} "
              f"stages {json.dumps({k: round(v, 2) for k, v in tm.items() if k.endswith('_ms')})}", flush=True)
        pair.close()
        A.normals = None if which != "c1" else A.normals
        B.normals = None if which != "c1" else B.normals
    for k, v in list(res.items())[:12]:
        print(k, v)


if __name__ == "__main__":
    main()
