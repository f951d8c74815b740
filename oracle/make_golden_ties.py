#!/usr/bin/env python
"""Golden values of the tie-averaged mode (beyond the reference: SURVEY 8(f)-4) -- TEST INFRASTRUCTURE.
Runs the brute-force restatement ``reference_port.tie_average`` on three committed fixtures (forced ties and
duplicates, a voxelised pair, a float pair) and stores the reductions the GPU must reproduce:
    python oracle/make_golden_ties.py        ->  tests/golden/tie_average.json
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import Golden  # noqa: E402
from oracle import reference_port as rp  # noqa: E402

out = {}
for name in ("ties", "vox_small", "float_small"):
    i = Golden(name).inputs()
    n = min(len(i["pts_a"]), len(i["pts_b"]))
    rng = np.random.default_rng(42)
    nrm = [rng.normal(0, 1, (n, 3)) for _ in range(2)]
    nrm = [v / np.linalg.norm(v, axis=1, keepdims=True) for v in nrm]
    col = [rng.integers(0, 256, (n, 3)).astype(np.float64) / 255.0 for _ in range(2)]
    o = rp.PairOracle(i["pts_a"][:n], i["pts_b"][:n], col[0], col[1], nrm[0], nrm[1])
    rec = {"n": int(n), "seed": 42}
    for is_left in (True, False):
        pe2, cd2 = rp.tie_average(o, is_left, "yuv")
        rec["left" if is_left else "right"] = {"sum_d2": float(pe2.sum()).hex(), "max_d2": float(pe2.max()).hex(),
                                              "color_sum": [float(x).hex() for x in cd2.sum(0)],
                                              "color_max": [float(x).hex() for x in cd2.max(0)]}
    out[name] = rec
with open(os.path.join(ROOT, "tests", "golden", "tie_average.json"), "w") as f:
    json.dump(out, f, indent=1)
print("wrote tests/golden/tie_average.json")
