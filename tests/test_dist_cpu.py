"""CPU tier: the multi-GPU host logic over gloo (world_size 2 and 3): query slices cover every
point exactly once, partial records from all ranks fold to the single-process result (integer
sums exactly, float sums to rounding, maxima exactly), the self-NN min/max and the disjoint
normal-buffer combine."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from open_pcc_metric_b200 import distributed as D


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_partial(rng, lo, hi, d2_all, pe_all, col_all):
    """What pccm_pair_eval would return for queries [lo, hi): reduced on this 'rank'."""
    d2, pe, col = d2_all[lo:hi], pe_all[lo:hi], col_all[lo:hi]
    return dict(sum_u64=int(d2.sum()), d2_valid=True, sum_d1=float(d2.sum()), max_d1=float(d2.max()),
                sum_d2=float(pe.sum()), max_d2=float(pe.max()), csum=col.sum(0), cmax=col.max(0))


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)              # same data on every rank (replicated clouds)
        dirs = []
        for d in range(2):
            m = n + 7 * d
            dirs.append((rng.integers(0, 1 << 20, m).astype(np.uint64), rng.random(m), rng.random((m, 3))))
        vals = []
        for d2, pe, col in dirs:
            lo, hi = D.slice_range(len(d2), rank, world)
            vals.append(_fake_partial(rng, lo, hi, d2, pe, col))
        out = D.exchange_partials(vals, world, None, "cpu")
        mn, mx = D.exchange_minmax(float(rank + 1), float(10 * (rank + 1)), None, "cpu")
        buf = torch.zeros((n, 3), dtype=torch.float64)
        lo, hi = D.slice_range(n, rank, world)
        full = torch.from_numpy(np.random.default_rng(1).normal(size=(n, 3)))
        buf[lo:hi] = full[lo:hi]
        D.combine_disjoint(buf)
        ok = True
        for (d2, pe, col), o in zip(dirs, out):
            ok &= o["sum_u64"] == int(d2.sum())
            ok &= o["max_d1"] == float(d2.max()) and o["max_d2"] == float(pe.max())
            ok &= bool(np.isclose(o["sum_d2"], pe.sum(), rtol=1e-13))
            ok &= bool(np.allclose(o["csum"], col.sum(0), rtol=1e-13)) and bool(np.array_equal(o["cmax"], col.max(0)))
            ok &= o["d2_valid"] is True
        ok &= (mn, mx) == (1.0, 10.0 * world)
        ok &= bool(torch.equal(buf, full))
        q.put((rank, bool(ok), [o["sum_d2"] for o in out]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_exchange_over_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 1001, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    sums = [s for _, _, s in res]
    assert all(s == sums[0] for s in sums)      # every rank folds in the same fixed order


def test_slice_ranges_partition_exactly():
    for n in (0, 1, 7, 1000, 1003976):
        for world in (1, 2, 3, 8):
            edges = [D.slice_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1
