#!/usr/bin/env python
"""How much of the query kernel is fixed cost?  Times pair_eval on (A, A) -- every query finds
d2 = 0 in its home pencil and prunes all rings -- with and without the D2 / colour epilogues."""
import os
import sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from open_pcc_metric_b200 import _native as N, synth  # noqa: E402

A, B = synth.synth_pair(10, 1_000_000, synth.BASE_SEED + 2, step=2, dedup=False, oversample=4)
ctx = N.Context(0)
YUV = np.array([[0.25, 0.5, 0.25], [1, 0, -1], [-0.5, 1, -0.5]])


def run(X, Y, flags, tag):
    a = ctx.cloud(X.points, X.colors, X.normals)
    b = ctx.cloud(Y.points, Y.colors, Y.normals)
    ctx.build_pair(a, b)
    for _ in range(3):
        ctx.pair_eval(a, b, flags, YUV)
    ctx.set_profiling(1)
    ctx.reset_timings()
    for _ in range(20):
        ctx.pair_eval(a, b, flags, YUV)
    t = ctx.timings()
    ctx.set_profiling(0)
    print(f"{tag:28s} query {t['query_ms'] / 20 * 1e3:7.1f} us per launch (2 x {len(X.points)} queries)")
    a.close(); b.close()


run(A, B, N.EVAL_D2 | N.EVAL_COLOR, "A vs B  D1+D2+colour")
run(A, B, 0, "A vs B  D1 only")
run(A, A, N.EVAL_D2 | N.EVAL_COLOR, "A vs A  D1+D2+colour")
run(A, A, 0, "A vs A  D1 only")
