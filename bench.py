#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json: symmetric D1+D2 NN
queries/sec; ms per 1M-point cloud pair).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (N=1): BASELINE.json configs[1] -- synthetic vox10 ~1M-point pair with RGB and
given normals; one STEP = one full symmetric evaluation of the pair from raw arrays:
statistics, pencil-grid index of both clouds (keys, radix sort, table, reorder), both NN
query passes fused with the D1 + D2 + YUV-colour epilogues and their reductions, the
result record read back, PSNRs on the host.
  value  whole-job queries/s with the raw arrays already resident in HBM (C ABI, DEVICE buffers)
  e2e    the same through the public drop-in API (CloudPair + MetricCalculator +
         transform_options) from PINNED HOST float64 arrays: host->device copies and the
         result read-back are inside the timed region; additionally includes the always-on
         MinSqrt/MaxSqrt self-NN pass of the reference's option expansion.
N>1: one process per GPU (torchrun); every rank evaluates its own pair of the same shape
(frames of a sequence sharded over GPUs); no collective on the data
path; value = queries of all ranks / max-over-ranks device time; scaling = weak.
``--mode partition`` instead splits ONE pair's queries over the ranks (replicated index,
NCCL exchange of the partial sums) -- strong scaling, reported when asked for.

--impl reference: the reference's own CPU structure (per-point Python loop over a KD-tree,
per-row np.dot / matmul loops) timed on this host by the oracle port (oracle/cpu_baseline.py),
on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "symmetric_nn_queries_per_sec_d1_d2_color"
UNIT = "queries/s"
WORKLOAD = "configs[1]: synthetic vox10 ~1M-pt pair + RGB + given normals, D1+D2+YUV colour PSNR"
ALG_BYTES_PER_QUERY = 44  # 12 B query xyz + 12 B search xyz + 12 B normal + 4 + 4 B colours
YUV = np.array([[0.25, 0.5, 0.25], [1, 0, -1], [-0.5, 1, -0.5]], dtype=np.float64)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--points", type=int, default=1_000_000, help="target points of the original cloud")
    ap.add_argument("--bits", type=int, default=10)
    ap.add_argument("--mode", default="frames", choices=["frames", "partition"])
    ap.add_argument("--cpu-sample", type=int, default=100_000, help="queries per direction timed on the CPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def make_pair(args, rank):
    from open_pcc_metric_b200 import synth
    # same-length quantised+jittered copy: the shape of the reference's own end-to-end test and the
    # only one for which its D2 is defined in both directions (quirk Q1)
    seed = synth.BASE_SEED + 2 + 1000 * rank
    return synth.synth_pair(args.bits, args.points, seed, step=2, dedup=False, oversample=4)


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons of one GPU sampled with NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            pass

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def psnr_summary(res, n_a, n_b, peak):
    """Host tail of a step: the PSNRs metric.py would report from the reduced sums."""
    out = {}
    with np.errstate(divide="ignore"):
        for d, n in ((0, n_a), (1, n_b)):
            r = res.dir[d]
            out[d] = (10 * np.log10(peak ** 2 / (r.sum_d1 / n)), 10 * np.log10(peak ** 2 / (r.sum_d2 / n)),
                      10 * np.log10(1.0 / (np.array(list(r.color_sum)) / n)))
    return out


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU structure, timed by the oracle port on this host."""
    if rank != 0:
        return
    from oracle import cpu_baseline as cb
    A, B = make_pair(args, 0)
    nq = len(A) + len(B)
    vals = []
    info = None
    for i in range(args.warmup + args.steps):
        sample = max(1000, args.cpu_sample // max(1, args.steps))
        info = cb.reference_structure(A, B, "yuv", True, sample=sample, seed=i)
        if i >= args.warmup:
            vals.append(info["queries_per_s"])
    v = float(np.mean(vals))
    best = cb.cpu_best(A, B, "yuv", True)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * nq / v, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "n_a": len(A), "n_b": len(B), "queries_per_step": nq},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": f"per step: KD-tree builds in full + {info['sample']} queries (both directions) through the "
                                   "reference's per-point loops, extrapolated linearly to the pair; single thread as in the reference",
                         "cpu_best": {"value": best["queries_per_s"], "cores": best["cores"],
                                      "what": "batched cKDTree.query(workers=-1) + vectorised numpy, full pair"}},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: open_pcc_metric_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        ge.build()           # no-op when the in-tree libpccm.so is current
    if world > 1:
        dist.barrier()       # nobody loads the library while rank 0 may still be linking it
    from open_pcc_metric_b200 import _native as N
    from open_pcc_metric_b200.calculator import MetricCalculator
    from open_pcc_metric_b200.cloud_pair import CloudPair
    from open_pcc_metric_b200.options import CalculateOptions, transform_options
    from open_pcc_metric_b200.synth import Cloud

    partition = args.mode == "partition" and world > 1
    A, B = make_pair(args, 0 if partition else rank)
    n_a, n_b = len(A), len(B)
    nq = n_a + n_b
    peak = float((1 << args.bits) - 1)

    stream = torch.cuda.Stream(device=dev)
    ctx = N.Context(local_rank, stream.cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    # raw inputs resident in HBM, in the reference's own layout (float64 points / normals, uchar colours)
    def to_dev(c):
        return (torch.from_numpy(c.points).to(dev), torch.from_numpy(np.rint(c.colors * 255).astype(np.uint8)).to(dev),
                torch.from_numpy(c.normals).to(dev))
    dA, dB = to_dev(A), to_dev(B)
    # pinned host copies for the end-to-end arm (float64 everywhere, as Open3D would hand them over)
    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True)
        t.numpy()[...] = a
        return t.numpy()
    hA = Cloud(pinned(A.points), pinned(A.colors), pinned(A.normals))
    hB = Cloud(pinned(B.points), pinned(B.colors), pinned(B.normals))
    h2d = sum(x.nbytes for c in (hA, hB) for x in (c.points, c.colors, c.normals))

    # the same clouds in the compact form a voxelised PLY file holds (uint16 coordinates, uchar colours,
    # float32 normals): accepted by the same public API, 3.4x fewer bytes over PCIe (informational arm)
    def pinned_as(a, dt):
        import torch as _t
        t = _t.empty(a.shape, dtype={np.uint16: _t.uint16, np.uint8: _t.uint8, np.float32: _t.float32}[dt], pin_memory=True)
        out = t.numpy()
        out[...] = a.astype(dt)
        return out
    cA = Cloud(pinned_as(A.points, np.uint16), pinned_as(np.rint(A.colors * 255), np.uint8), pinned_as(A.normals, np.float32))
    cB = Cloud(pinned_as(B.points, np.uint16), pinned_as(np.rint(B.colors * 255), np.uint8), pinned_as(B.normals, np.float32))
    h2d_compact = sum(x.nbytes for c in (cA, cB) for x in (c.points, c.colors, c.normals))

    sl = (rank, world) if partition else (0, 1)

    def step_device():
        a = ctx.cloud(*dA)
        b = ctx.cloud(*dB)
        ctx.build_pair(a, b)
        res = ctx.pair_eval(a, b, N.EVAL_D2 | N.EVAL_COLOR, YUV, 1.0, N.NORMALS_BY_QUERY_INDEX, sl[0], sl[1])
        a.close()
        b.close()
        return res

    opts = CalculateOptions(color="yuv", hausdorff=False, point_to_plane=True)

    def step_e2e(a=None, b=None):
        pair = CloudPair(a or hA, b or hB, ctx=ctx, peak="resolution", resolution_bits=args.bits,
                         rank=sl[0], world=sl[1])
        out = MetricCalculator(pair).calculate(transform_options(opts)).as_dict()
        pair.close()
        return out

    def timed(fn, steps, warmup, profile_level):
        with torch.cuda.stream(stream):
            for _ in range(warmup):
                flush.zero_()
                fn()
            ctx.set_profiling(profile_level)
            ctx.reset_timings()
            evs = []
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            last = None
            for _ in range(steps):
                flush.zero_()                      # L2 flush between iterations, outside the event pair
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                last = fn()
                e1.record(stream)
                evs.append((e0, e1))
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            wall = time.perf_counter() - t0
            ms = sum(a.elapsed_time(b) for a, b in evs)
            tm = ctx.timings()
            ctx.set_profiling(0)
        return ms, wall, tm, last

    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_dev, wall_dev, tm, res = timed(step_device, args.steps, args.warmup, 1)
    clocks = sampler.stop()

    if world > 1:
        t = torch.tensor([ms_dev], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dev_max = float(t.item())
    else:
        ms_dev_max = ms_dev
    total_queries = nq * args.steps * (1 if partition else world)
    value = total_queries / (ms_dev_max * 1e-3)

    # query-kernel roofline (one launch per step covers A->B and B->A), measured live with CUDA events
    q_launches = max(1, tm["query_launches"])
    brick = tm.get("vox_epilogue_ms", 0) > 0
    # brick path: the query stage is three launches (search of one lane per voxel; brick-ring search of the voxels
    # it left undecided; per-point epilogue); query_ms brackets the whole stage, the split is reported beside it
    q_ms_avg = tm["query_ms"] / q_launches
    q_split = {"vx_search_kernel": tm["vox_search_ms"] / q_launches,
               "vx_general_kernel": tm.get("vox_tail_ms", 0) / q_launches,
               "vx_epilogue_kernel": tm.get("vox_epilogue_ms", 0) / q_launches} if brick else None
    queries_per_launch = nq / (world if partition else 1)
    achieved = ALG_BYTES_PER_QUERY * queries_per_launch / (q_ms_avg * 1e-3) / 1e9
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak_gbs, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    else:
        peak_gbs, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    traffic = None
    tp = os.path.join(ROOT, "profiles", "query_kernel_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("dram_bytes_per_launch")

    # per-stage device time (separate short loop with every stage bracketed by events; informational)
    _, _, tm2, _ = timed(step_device, 5, 1, 2)
    stages = {k: round(v / 5, 5) for k, v in tm2.items() if k.endswith("_ms")}

    e2e = None
    if not args.no_e2e:
        ms_e2e, _, _, out = timed(step_e2e, max(3, args.steps // 2), 3, 0)
        n_e2e = max(3, args.steps // 2)
        if world > 1:
            t = torch.tensor([ms_e2e], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_e2e = float(t.item())
        _, _, tm3, _ = timed(step_e2e, 5, 1, 2)
        e2e_stages = {k: round(v / 5, 5) for k, v in tm3.items() if k.endswith("_ms")}
        e2e = {"value": nq * n_e2e * (1 if partition else world) / (ms_e2e * 1e-3), "unit": UNIT,
               "stage_ms_per_step": e2e_stages,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(2 * 96 + 16),
               "ms_per_step": ms_e2e / n_e2e,
               "api": "CloudPair(host float64 arrays) + MetricCalculator.calculate(transform_options(color=yuv, point_to_plane))"}
        ms_c, _, _, out_c = timed(lambda: step_e2e(cA, cB), max(3, args.steps // 4), 3, 0)
        n_c = max(3, args.steps // 4)
        same = all(np.array_equal(np.asarray(out[k]), np.asarray(out_c[k])) for k in out if k[0] in ("GeoMSE", "ColorMSE") and k[-1] is not True)
        e2e["compact_inputs"] = {"value": nq * n_c / (ms_c * 1e-3), "ms_per_step": ms_c / n_c, "h2d_bytes_per_step": int(h2d_compact),
                                 "what": "same API, clouds held as uint16 / uchar / float32 (float32 normals change D2 only)",
                                 "d1_and_colour_identical_to_float64_inputs": bool(same)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import cpu_baseline as cb
        info = cb.reference_structure(A, B, "yuv", True, sample=args.cpu_sample)
        best = cb.cpu_best(A, B, "yuv", True)
        cpu = {"value": info["queries_per_s"], "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"KD-tree builds in full + {info['sample']} queries through the reference's per-point loops "
                         f"({info['per_query_us']:.1f} us/query), extrapolated linearly to the {nq}-query pair; "
                         "single thread as in the reference",
               "cpu_best": {"value": best["queries_per_s"], "cores": best["cores"],
                            "what": "batched cKDTree.query(workers=-1) + vectorised numpy, full pair"}}

    if rank == 0:
        ps = psnr_summary(res, n_a, n_b, peak) if not partition else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev_max / args.steps, "higher_is_better": True,
            "scaling": "strong" if partition else "weak", "vs_baseline": None, "dtype": "u32 distances / f64 epilogues",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "n_a": n_a, "n_b": n_b, "queries_per_step_per_gpu": nq if not partition else nq // world,
                       "coordinate_kind": "int (vox%d)" % args.bits, "normals_mode": "reference (by query index)",
                       "peak": "resolution", "parallelism": ("query slices of one pair" if partition else "one pair per GPU"),
                       "l2": "flushed between iterations (256 MiB write)", "ms_per_1M_point_pair": ms_dev_max / args.steps},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs,
                         "traffic": traffic, "kernel": "query stage = vx_search_kernel + vx_general_kernel + vx_epilogue_kernel" if brick else "pair_query_kernel<KInt>",
                         "launch_ms_by_kernel": q_split,
                         # the two kernels of the stage against their own share of the algorithmic traffic (SURVEY 8(d):
                         # 24 B/query for the search -- query + search coordinates; 20 B/query for the epilogue -- normal + two colours)
                         "by_kernel": None if not brick else {
                             k: {"alg_bytes_per_query": ab, "achieved": ab * queries_per_launch / (q_split[k] * 1e-3) / 1e9,
                                 "frac": ab * queries_per_launch / (q_split[k] * 1e-3) / 1e9 / peak_gbs}
                             for k, ab in (("vx_search_kernel", 24), ("vx_epilogue_kernel", 20)) if q_split[k] > 0},
                         "peak_source": peak_src,
                         "alg_bytes_per_query": ALG_BYTES_PER_QUERY, "queries_per_launch": queries_per_launch,
                         "avg_launch_ms": q_ms_avg},
            "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks,
            "gpu_launches": int(tm["total_launches"]), "library_launches": int(tm["library_launches"]),
            "wall_s_timed_region": wall_dev, "stage_ms_per_step": stages,
            "brick_path": {k: int(tm[k]) for k in ("vox_undecided", "vox_far", "vox_tail") if k in tm},
            "check": None if ps is None else {"d1_psnr_left": float(ps[0][0]), "d2_psnr_left": float(ps[0][1]),
                                              "y_psnr_left": float(ps[0][2][0])},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
