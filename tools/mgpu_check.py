#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun, one process per GPU):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py
Every rank evaluates slice rank/world of the same pair (replicated indices, normals estimated in
slices and combined by all-reduce, partial records exchanged by all-gather); the result must match
a single-process evaluation: integer D1 family bit for bit, float sums to 1e-12."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from open_pcc_metric_b200 import _native as N, synth  # noqa: E402
from open_pcc_metric_b200.calculator import MetricCalculator  # noqa: E402
from open_pcc_metric_b200.cloud_pair import CloudPair  # noqa: E402
from open_pcc_metric_b200.options import CalculateOptions, transform_options  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = N.Context(local)
    A, B = synth.synth_pair(9, 200_000, 77, with_normals=False, dedup=False, oversample=4)
    opts = CalculateOptions(color="ycc", hausdorff=True, point_to_plane=True)

    def run(r, w):
        a = synth.Cloud(A.points, A.colors, None)
        b = synth.Cloud(B.points, B.colors, None)
        pair = CloudPair(a, b, ctx=ctx, peak="resolution", resolution_bits=9, rank=r, world=w)
        out = MetricCalculator(pair).calculate(transform_options(opts)).as_dict()
        nrm = pair.get_normals(0).copy()
        pair.close()
        return out, nrm

    sharded, nrm_s = run(rank, world)
    ok = True
    if rank == 0:
        single, nrm_1 = run(0, 1)
        assert set(single) == set(sharded)
        for k, v in single.items():
            g, w = np.asarray(sharded[k], float), np.asarray(v, float)
            name = k[0] if k[0] != "SymmetricMetric" else k[1]
            p2p_or_color = any(x is True for x in k[2:4]) or any(isinstance(x, str) and x in ("ycc",) for x in k)
            if name in ("MinSqrtDistance", "MaxSqrtDistance") or (name.startswith("Geo") and not p2p_or_color):
                good = np.array_equal(g, w)
            else:
                good = np.allclose(g, w, rtol=1e-12, atol=1e-30)
            if not good:
                ok = False
                print("MISMATCH", k, g, w, flush=True)
        if not np.array_equal(nrm_s, nrm_1):
            ok = False
            print("normals differ", np.abs(nrm_s - nrm_1).max(), flush=True)
        print(f"[mgpu_check] world={world} {len(single)} metrics: {'OK' if ok else 'FAILED'}", flush=True)
    flag = torch.tensor([0 if ok else 1], device=f"cuda:{local}")
    dist.all_reduce(flag)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(int(flag.item() != 0))


if __name__ == "__main__":
    main()
