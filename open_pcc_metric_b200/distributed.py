"""Multi-GPU plumbing: one process per GPU, query slices
per rank, and the only data that crosses NVLink -- the tiny per-direction partial
records (C3), the self-NN min/max, and (when normals are estimated) one all-reduce
of disjointly filled normal buffers (C2).  torch.distributed does the transport
(NCCL on GPUs; the same code runs over gloo on CPU tensors in the tests).
"""
from __future__ import annotations

import numpy as np

SUM_SLOTS = (0, 2, 4, 5, 6)      # sum_d1, sum_d2, colour sums
MAX_SLOTS = (1, 3, 7, 8, 9)      # max_d1, max_d2, colour maxima


def slice_range(n: int, rank: int, world: int):
    """Contiguous equal-count slice of a cloud's sorted order (same rule as pccm_pair_eval)."""
    return n * rank // world, n * (rank + 1) // world


def exchange_partials(vals, world: int, group=None, device="cpu"):
    """vals: per direction dict(sum_u64, d2_valid, sum_d1, max_d1, sum_d2, max_d2, csum[3], cmax[3][, n]).
    ONE all_gather of every rank's record (the 64-bit integer sum travels as two exact 32-bit halves) and a
    reduction on the host: integer sums exactly (Python ints), float sums in fixed rank order (deterministic
    for a given world)."""
    import torch
    import torch.distributed as dist
    rows = []
    for v in vals:
        u = int(v["sum_u64"])
        rows.append([v["sum_d1"], v["max_d1"], v["sum_d2"], v["max_d2"], *v["csum"], *v["cmax"],
                     float(u >> 32), float(u & 0xffffffff), float(bool(v["d2_valid"])), float(v.get("n", 0))])
    fl = torch.tensor(rows, dtype=torch.float64, device=device)
    try:                                         # ONE output tensor: no per-rank list to copy in and out of
        gathered = torch.empty((world * fl.shape[0], fl.shape[1]), dtype=torch.float64, device=device)   # (concatenated form: every backend)
        dist.all_gather_into_tensor(gathered, fl, group=group)
        gathered = gathered.view(world, fl.shape[0], fl.shape[1])
    except (RuntimeError, NotImplementedError):  # a backend without the tensor form
        parts = [torch.empty_like(fl) for _ in range(world)]
        dist.all_gather(parts, fl, group=group)
        gathered = torch.stack(parts)
    rows_all = gathered.cpu().tolist()           # [world][ndir][14] Python floats: the fold below is ~10 us at world 8
    out = []                                     # (numpy scalars made it ~110 us -- inside every step of a split pair)
    for d, v in enumerate(vals):
        r = dict(v)
        u, valid, n = 0, True, 0
        acc = [0.0] * 10
        for j in MAX_SLOTS:
            acc[j] = float("-inf")
        for w in range(world):                   # fixed rank order: the float sums are reproducible for a given world
            row = rows_all[w][d]
            for j in SUM_SLOTS:
                acc[j] = acc[j] + row[j]
            for j in MAX_SLOTS:
                if row[j] > acc[j]:
                    acc[j] = row[j]
            u += (int(row[10]) << 32) + int(row[11])
            valid = valid and row[12] != 0.0
            n += int(row[13])
        r["sum_u64"] = u
        r["d2_valid"] = valid
        if "n" in v:
            r["n"] = n
        r["sum_d1"], r["max_d1"], r["sum_d2"], r["max_d2"] = acc[0], acc[1], acc[2], acc[3]
        r["csum"], r["cmax"] = np.array(acc[4:7]), np.array(acc[7:10])
        out.append(r)
    return out


def exchange_minmax(mn: float, mx: float, group=None, device="cpu"):
    import torch
    import torch.distributed as dist
    t = torch.tensor([-mn, mx], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    t = t.cpu().numpy()
    return -float(t[0]), float(t[1])


def combine_disjoint(t, group=None):
    """Every element was written by exactly one rank and is zero elsewhere: x + 0 == x."""
    import torch.distributed as dist
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t
