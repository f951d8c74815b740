#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json: symmetric D1+D2 NN
queries/sec at 1/2/4/8 B200; ms per 1M-point cloud pair).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config 2|split|3|4|5] [--mode frames]

One STEP = one full symmetric evaluation from raw arrays: statistics + packed coordinates,
occupancy-brick index of both clouds (directory, occupancy rows, ranks, voxel records), both NN passes
(bit-scan search per voxel, per-point epilogue with D1 / D2 / colour) and their reductions, the result
record read back.  Nothing is cached between steps.

Workloads (``config.workload`` names the one that ran):
  --config 2      BASELINE configs[1], the default at N=1: synthetic vox10 ~1M-point pair with RGB and
                  given normals, D1 + D2 + YUV.
  --config split  the default at N>1 (one process per GPU, torchrun): ONE vox12 ~10M-point pair with RGB and
                  given normals, D1 + D2 + YUV, split over the ranks by slabs of z inside the library
                  (pccm_ctx_set_shard: every rank holds both clouds, indexes its slab + halo, evaluates the
                  queries of its slab); the ranks' partial sums are exchanged with NCCL (all_gather) INSIDE the
                  timed step.  scaling = strong; the line also carries the same pair evaluated on one GPU
                  (``strong_scaling_reference``) so that the speed-up can be read from one line.
  --mode frames   N>1 alternative (weak scaling): every rank evaluates its own config-2 pair (frames of a
                  sequence sharded over GPUs), no collective on the data path.
  --config 3      BASELINE configs[2]: vox12 ~4M-point pair without normals: k=30 kNN+PCA normals for both
                  clouds, then D1 / D2 / Hausdorff and the boundary distances.
  --config 4      BASELINE configs[3]: frames of ~800k-point vox10 clouds through the pipelined sequence API
                  (evaluate_sequence), frames sharded over the ranks.
  --config 5      BASELINE configs[4]: float32 LiDAR-style pair (default 50M points), normals estimated,
                  D1 / D2 / Hausdorff.

  value  whole-job queries/s with the raw arrays already resident in HBM (C ABI, DEVICE buffers)
  e2e    the same through the public drop-in API (CloudPair + MetricCalculator + transform_options) from
         PINNED HOST float64 arrays: host->device copies and the result read-back are inside the timed
         region; additionally includes the always-on MinSqrt/MaxSqrt self-NN pass of the reference's option
         expansion.  Split pair (N>1): every rank hands BOTH clouds to CloudPair(rank=, world=) -- the split's
         contract -- so h2d_bytes_per_step is per rank; the NCCL exchanges are inside the step.

--impl reference: the reference's CPU path timed on this host by the oracle port (oracle/cpu_baseline.py): per
step ONE FULL evaluation of the same pair in batched form on all cores (cKDTree.query(workers=-1) + whole-array
numpy) -- measured, not extrapolated; the reference's own single-threaded per-point loop structure is reported
beside it (in full on configs[0], sampled on the workload).
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
YUV = np.array([[0.25, 0.5, 0.25], [1, 0, -1], [-0.5, 1, -0.5]], dtype=np.float64)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="auto", choices=["auto", "2", "split", "3", "4", "5"])
    ap.add_argument("--mode", default="auto", choices=["auto", "frames", "partition"])
    ap.add_argument("--points", type=int, default=None, help="target points of the original cloud (default: the config's size)")
    ap.add_argument("--cpu-sample", type=int, default=100_000, help="queries per direction timed through the reference's per-point loops")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--shard-of", default=None, help="R,W: time what rank R of W would do on a split pair, on one GPU, without the exchange (tuning aid)")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------------------------------
class Workload:
    """What one step evaluates.  `config` is the dict both arms print (identical keys and values)."""

    def __init__(self, args, world):
        from open_pcc_metric_b200 import synth
        cfg = args.config
        if cfg == "auto":
            cfg = "2" if (world == 1 or args.mode == "frames") else "split"
        if args.mode == "partition":
            cfg = "split"
        self.cfg = cfg
        self.frames_mode = world > 1 and args.mode == "frames" and cfg == "2"
        self.split = world > 1 and cfg == "split"
        self.color = "yuv" if cfg in ("2", "split", "4") else None
        self.estimate_normals = cfg in ("3", "5")
        self.hausdorff = cfg in ("3", "5")
        self.alg_bytes = 44 if self.color else 36        # 24 B coordinates + 12 B normal (+ 4 + 4 B colours)
        if cfg == "2":
            self.bits, self.points = 10, args.points or 1_000_000
            self.name = "configs[1]: synthetic vox10 ~1M-pt pair + RGB + given normals, D1+D2+YUV colour PSNR"
            self.gen = lambda rank: synth.synth_pair(self.bits, self.points, synth.BASE_SEED + 2 + 1000 * (rank if self.frames_mode else 0),
                                                     step=2, dedup=False, oversample=4)
        elif cfg == "split":
            self.bits, self.points = 12, args.points or 10_500_000
            self.name = "north_star: ONE synthetic vox12 >=10M-pt pair + RGB + given normals, D1+D2+YUV, z-slab split over the GPUs"
            self.gen = lambda rank: self._vox12(synth, True, oversample=2, tol=0.04)
        elif cfg == "3":
            self.bits, self.points = 12, args.points or 4_000_000
            self.name = "configs[2]: synthetic vox12 ~4M-pt pair without normals: kNN+PCA normals, D1/D2/Hausdorff"
            self.gen = lambda rank: self._vox12(synth, False)
        elif cfg == "4":
            self.bits, self.points = 10, args.points or 800_000
            self.name = "configs[3]: frames of ~800k-pt vox10 clouds (sequence evaluation), D1+D2+YUV, frames sharded over the GPUs"
            self.gen = lambda rank: synth.synth_pair(self.bits, self.points, synth.BASE_SEED + 4 + 1000 * rank, step=2, dedup=False, oversample=4)
        else:
            self.bits, self.points = None, args.points or 50_000_000
            self.name = "configs[4]: synthetic float32 LiDAR-style pair, normals estimated, D1/D2/Hausdorff"
            self.gen = lambda rank: self._lidar(synth)
        self.peak = float((1 << self.bits) - 1) if self.bits else None

    def _vox12(self, synth, attrs, oversample=3, tol=0.02):
        # (generating 10M voxelised surface points takes a minute of numpy: keep them for later runs on the same box)
        cache = os.path.join("/tmp", f"pccm_bench_vox12_{self.points}_{int(attrs)}_{oversample}.npz")
        A = None
        if os.path.exists(cache):
            try:
                z = np.load(cache)
                A = synth.Cloud(z["p"], z["c"] if attrs else None, z["n"] if attrs else None)
            except Exception:                     # (a truncated file of an interrupted run: generate again)
                A = None
        if A is None:
            A = synth.synth_vox(12, self.points, synth.BASE_SEED + 3, with_colors=attrs, with_normals=attrs, oversample=oversample, tol=tol)
            # the ranks of one job all arrive here together: rank 0 alone keeps the arrays (0.7 GB for the 10 M pair), in a
            # temporary file renamed over the cache name (atomic: a reader sees a complete file or none)
            if int(os.environ.get("RANK", "0")) == 0:
                tmp = f"{cache}.{os.getpid()}.tmp.npz"
                try:
                    np.savez(tmp, p=A.points, c=A.colors if attrs else np.zeros(0), n=A.normals if attrs else np.zeros(0))
                    os.replace(tmp, cache)
                except OSError:
                    try:
                        os.remove(tmp)
                    except OSError:
                        pass
        return A, synth.degrade(A, 2, synth.BASE_SEED + 3, 12, dedup=False)

    def _lidar(self, synth):
        # one point of B per point of A (the reference's D2 needs |search| >= |query| in both directions, quirk Q1):
        # both clouds have the stated size
        return synth.synth_lidar(self.points, synth.BASE_SEED + 5, dedup=False)

    def config(self, n_a, n_b, world, distinct_b=None):
        return {
            "workload": self.name, "n_a": int(n_a), "n_b": int(n_b), "queries_per_step": int(n_a + n_b),
            "degraded_cloud": "one point per input point (dedup=False: quirk Q1 -- the reference's D2 is only defined when the search "
                              "cloud is at least as long as the query cloud)" if self.bits else "one point per input point: 2 cm lattice + N(0, 5 mm) jitter",
            "distinct_voxels_b": None if distinct_b is None else int(distinct_b),
            "coordinate_kind": ("int (vox%d)" % self.bits) if self.bits else "float32",
            "inputs": "float64 coordinates and normals, uchar colours (device arm); float64 everywhere (e2e arm)" if self.bits else "float32 coordinates",
            "normals": "estimated (k=30, PCA)" if self.estimate_normals else "given",
            "normals_mode": "reference (by query index)", "peak": "resolution" if self.bits else "aabb_diag",
            "parallelism": ("one pair per GPU" if self.frames_mode else ("one pair, z slabs" if self.split else ("frames round-robin" if self.cfg == "4" else "single GPU"))),
            "l2": "flushed between iterations (256 MiB write)",
        }


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


def psnr_summary(fused, n_a, n_b, peak):
    """Host tail of a step: the PSNRs metric.py would report from the reduced sums."""
    out = {}
    with np.errstate(divide="ignore", invalid="ignore"):
        for d, n in ((0, n_a), (1, n_b)):
            r = fused[d]
            out[d] = (10 * np.log10(peak ** 2 / (r["sum_d1"] / n)), 10 * np.log10(peak ** 2 / (r["sum_d2"] / n)),
                      10 * np.log10(1.0 / (np.array(r["csum"]) / n)))
    return out


# ---------------------------------------------------------------------------------------------------
# reference arm
# ---------------------------------------------------------------------------------------------------
def cpu_reference(W, A, B, sample, prebuilt=None):
    """The reference's per-point loop structure (1 thread): in full on configs[0], sampled on the workload."""
    from oracle import cpu_baseline as cb
    from open_pcc_metric_b200 import synth
    a1, b1 = synth.synth_pair(10, 100_000, synth.BASE_SEED + 1, with_colors=False)
    t0 = time.perf_counter()
    full = cb.reference_structure(a1, b1, None, False, sample=max(len(a1), len(b1)))
    full_s = time.perf_counter() - t0
    info = cb.reference_structure(A, B, W.color, True, sample=sample, prebuilt=prebuilt)
    return {
        "configs0_full": {"queries": len(a1) + len(b1), "seconds": full_s, "queries_per_s": (len(a1) + len(b1)) / full["measured_s"],
                          "what": "configs[0] (vox10 ~100k pair, D1) through the reference's per-point loops, every query, 1 thread: measured"},
        "workload_sampled": {"queries_per_s": info["queries_per_s"], "per_query_us": info["per_query_us"], "sample": info["sample"],
                             "what": "KD-tree builds in full + sampled queries through the per-point loops, extrapolated linearly; 1 thread"},
    }


def run_reference(args, rank, world):
    """--impl reference: per step one FULL evaluation of the pair on the host cores (batched port)."""
    if rank != 0:
        return
    from oracle import cpu_baseline as cb
    W = Workload(args, world)
    A, B = W.gen(0)
    if W.estimate_normals:
        raise SystemExit("--impl reference supports the colour / given-normals workloads (configs 2, split, 4)")
    nq = len(A) + len(B)
    steps = args.steps or 5
    vals = []
    # Pairs up to a few million points: every step is one FULL evaluation (about a second).  The 10 M-point pair of the
    # multi-GPU runs takes half a minute per evaluation on 16 cores -- K + W of them would run for a quarter of an hour --
    # so a step is a BOUNDED SAMPLE of it: the two KD-trees are built once (timed), every step evaluates the first million
    # queries of each direction against them on all cores, and the pair's time is build + sample x (queries / sample).
    big = nq > 6_000_000
    prebuilt = cb.build_trees(A, B) if big else None
    for i in range(args.warmup + steps):
        best = cb.cpu_best(A, B, W.color, True, 1_000_000, prebuilt) if big else cb.cpu_best(A, B, W.color, True)
        if i >= args.warmup:
            vals.append(best["queries_per_s"])
    v = float(np.mean(vals))
    loops = cpu_reference(W, A, B, max(1000, args.cpu_sample // 10), prebuilt)
    sample_text = ("every step evaluates the FULL pair (no extrapolation): KD-tree builds + cKDTree.query(workers=-1) on all cores + "
                   "whole-array numpy D1 / D2 / colour; the reference itself runs these per point from one Python thread -- see reference_structure")
    if big:
        sample_text = (f"bounded sample of the {nq}-query pair: both KD-trees built in full once ({prebuilt[1]:.1f} s, counted in every step), every step "
                       f"evaluates the first 1,000,000 queries of each direction against them (cKDTree.query(workers=-1) on all cores + whole-array numpy "
                       f"D1 / D2 / colour) and is scaled to the pair: build + sample time x (queries / sample); the reference itself runs these per point "
                       f"from one Python thread -- see reference_structure")
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * nq / v, "higher_is_better": True,
        "scaling": "strong" if W.split else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": W.config(len(A), len(B), world, distinct_voxels(B.points) if W.bits else None),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cb.host_cores(), "kind": "port", "sample": sample_text, "sampled": bool(big),
                         "reference_structure": loops},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def distinct_voxels(pts):
    p = np.asarray(pts)
    key = (p[:, 0].astype(np.int64) << 42) | (p[:, 1].astype(np.int64) << 21) | p[:, 2].astype(np.int64)
    return int(np.unique(key).size)


# ---------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------
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
    from open_pcc_metric_b200 import distributed as D
    from open_pcc_metric_b200.calculator import MetricCalculator
    from open_pcc_metric_b200.cloud_pair import CloudPair
    from open_pcc_metric_b200.options import CalculateOptions, transform_options
    from open_pcc_metric_b200.synth import Cloud

    W = Workload(args, world)
    steps = args.steps or (400 if W.cfg == "2" else (50 if W.cfg in ("split", "4") else 10))
    stream = torch.cuda.Stream(device=dev)
    ctx = N.Context(local_rank, stream.cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    # ---- inputs: every rank needs the same pair when it is split; rank 0 generates, NCCL broadcasts
    def to_dev(c):
        pts = torch.from_numpy(np.ascontiguousarray(c.points)).to(dev)
        col = None if c.colors is None else torch.from_numpy(np.rint(c.colors * 255).astype(np.uint8)).to(dev)
        nrm = None if c.normals is None else torch.from_numpy(c.normals).to(dev)
        return [pts, col, nrm]

    A = B = None
    if W.split:
        if rank == 0:
            A, B = W.gen(0)
            dA, dB = to_dev(A), to_dev(B)
            meta = [len(A), len(B), distinct_voxels(B.points)]
        else:
            meta = [0, 0, 0]
        mt = torch.tensor(meta, dtype=torch.int64, device=dev)
        dist.broadcast(mt, 0)
        n_a, n_b, distinct_b = (int(x) for x in mt.tolist())
        if rank != 0:
            dA = [torch.empty((n_a, 3), dtype=torch.float64, device=dev), torch.empty((n_a, 3), dtype=torch.uint8, device=dev),
                  torch.empty((n_a, 3), dtype=torch.float64, device=dev)]
            dB = [torch.empty((n_b, 3), dtype=torch.float64, device=dev), torch.empty((n_b, 3), dtype=torch.uint8, device=dev),
                  torch.empty((n_b, 3), dtype=torch.float64, device=dev)]
        for t in dA + dB:
            dist.broadcast(t, 0)
    else:
        A, B = W.gen(rank)
        dA, dB = to_dev(A), to_dev(B)
        n_a, n_b = len(A), len(B)
        distinct_b = distinct_voxels(B.points) if W.bits else None
    nq = n_a + n_b

    flags = (N.EVAL_D2) | (N.EVAL_COLOR if W.color else 0)
    emulate = tuple(int(x) for x in args.shard_of.split(",")) if args.shard_of else None
    devstr = f"cuda:{local_rank}"

    def fused_dict(res):
        return [dict(n=int(r.n), sum_u64=int(r.sum_d1_u64), d2_valid=bool(r.d2_valid), sum_d1=float(r.sum_d1), max_d1=float(r.max_d1),
                     sum_d2=float(r.sum_d2), max_d2=float(r.max_d2), csum=list(r.color_sum), cmax=list(r.color_max)) for r in (res.dir[0], res.dir[1])]

    # ---- the step: C ABI, device-resident raw arrays
    def step_device(shard=None):
        sh = shard if shard is not None else ((rank, world) if W.split else (emulate or (0, 1)))
        if sh[1] > 1:
            ctx.set_shard(*sh)
        try:
            a = ctx.cloud(*dA)
            b = ctx.cloud(*dB)
            ctx.build_pair(a, b)
        finally:
            if sh[1] > 1:
                ctx.set_shard(0, 1)
        if W.estimate_normals:
            a.estimate_normals(30)
            b.estimate_normals(30)
        res = ctx.pair_eval(a, b, flags, YUV if W.color else None, 1.0, N.NORMALS_BY_QUERY_INDEX, 0, 1)
        out = fused_dict(res)
        if W.hausdorff:
            out[0]["self_nn"] = a.self_nn_minmax()[:2]
        if sh[1] > 1 and shard is None and world > 1:
            out = D.exchange_partials(out, world, None, devstr)      # NCCL all_gather of the partial records, inside the step
        a.close()
        b.close()
        return out

    # ---- the same through the drop-in surface from pinned host float64 arrays
    opts = CalculateOptions(color=W.color, hausdorff=W.hausdorff, point_to_plane=True)

    def step_e2e(a, b):
        pair = CloudPair(a, b, ctx=ctx, peak="resolution" if W.bits else "aabb_diag", resolution_bits=W.bits,
                         rank=rank if W.split else 0, world=world if W.split else 1)
        out = MetricCalculator(pair).calculate(transform_options(opts)).as_dict()
        pair.close()
        return out

    def timed(fn, nsteps, warmup, profile_level):
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
            for _ in range(nsteps):
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
            per_step = [a.elapsed_time(b) for a, b in evs]
            ms = sum(per_step)
            if os.environ.get("PCCM_BENCH_VERBOSE"):
                q = sorted(per_step)
                print(f"[timed] {getattr(fn, '__name__', 'step')}: n={len(q)} min={q[0]:.3f} median={q[len(q) // 2]:.3f} max={q[-1]:.3f} ms; "
                      f"slowest at steps {sorted(range(len(per_step)), key=lambda i: -per_step[i])[:5]}", file=sys.stderr, flush=True)
            tm = ctx.timings()
            ctx.set_profiling(0)
        return ms, wall, tm, last

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_dev, wall_dev, tm, res = timed(step_device, steps, args.warmup, 1)
    clocks = sampler.stop()
    ms_dev_max = max_over_ranks(ms_dev)
    per_rank_jobs = world if (W.frames_mode or W.cfg == "4") and world > 1 else 1
    total_queries = nq * steps * per_rank_jobs
    value = total_queries / (ms_dev_max * 1e-3)

    # per-stage device time (separate short loop with every stage and every kernel of the query stage bracketed by events;
    # the event pairs keep those kernels from overlapping their launches, so the timed region above brackets the stage only)
    _, _, tm2, _ = timed(step_device, 10, 1, 2)
    stages = {k: round(v / 10, 5) for k, v in tm2.items() if k.endswith("_ms")}

    # ---- roofline of the dominant kernel(s), measured live with CUDA events inside the library
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak_gbs, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    else:
        peak_gbs, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    q_launches = max(1, tm["query_launches"])
    brick = tm2.get("vox_epilogue_ms", 0) > 0
    if W.estimate_normals:
        # normal estimation dominates: k=30 self k-NN + PCA, 24 B / point (12 B read + 12 B normal written)
        k_ms = tm["knn_ms"] / steps
        units = nq
        achieved = 24 * units / (k_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs, "traffic": None,
                "kernel": "normals_int_kernel + knn_self_kernel (k=30 self k-NN + PCA) and the boundary-distance pass",
                "alg_bytes_per_point": 24, "points_per_step": units, "avg_ms_per_step": k_ms, "peak_source": peak_src,
                "note": "compute bound by design (counting selection / top-k lists), reported against the HBM roofline the task names"}
    else:
        q_ms_avg = tm["query_ms"] / q_launches
        q2 = max(1, tm2["query_launches"])
        q_split = {"vx_search_kernel": tm2["vox_search_ms"] / q2, "vx_general_kernel": tm2.get("vox_tail_ms", 0) / q2,
                   "vx_epilogue_kernel": tm2.get("vox_epilogue_ms", 0) / q2} if brick else None
        own = sum(r["n"] for r in step_device(shard=(rank, world))) if W.split else nq     # queries this rank's launches reduce
        achieved = W.alg_bytes * own / (q_ms_avg * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "query_kernel_traffic.json")
        if os.path.exists(tp) and W.cfg == "2" and not W.frames_mode:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs, "traffic": traffic,
                "kernel": "query stage = vx_search_kernel + vx_general_kernel + vx_epilogue_kernel" if brick else "pair_query_kernel",
                "launch_ms_by_kernel": q_split,          # (from the separate stage loop: events around every kernel)
                # the two kernels of the stage against their own share of the algorithmic traffic (24 B/query
                # for the search -- query + search coordinates; the rest for the epilogue -- normal + two colours)
                "by_kernel": None if not brick else {
                    k: {"alg_bytes_per_query": ab, "achieved": ab * own / (q_split[k] * 1e-3) / 1e9, "frac": ab * own / (q_split[k] * 1e-3) / 1e9 / peak_gbs}
                    for k, ab in (("vx_search_kernel", 24), ("vx_epilogue_kernel", W.alg_bytes - 24)) if q_split[k] > 0},
                "peak_source": peak_src, "alg_bytes_per_query": W.alg_bytes, "queries_per_launch": own, "avg_launch_ms": q_ms_avg}


    # strong scaling: the same pair, unsplit, on ONE GPU (rank 0, outside the timed region)
    strong_ref = None
    if W.split:
        if rank == 0:
            with torch.cuda.stream(stream):
                step_device(shard=(0, 1))
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for _ in range(3):
                    r1 = step_device(shard=(0, 1))
                e1.record(stream)
                torch.cuda.synchronize()
                ms1 = e0.elapsed_time(e1) / 3
            same = all(r1[d]["sum_u64"] == res[d]["sum_u64"] and r1[d]["max_d1"] == res[d]["max_d1"] and
                       np.isclose(r1[d]["sum_d2"], res[d]["sum_d2"], rtol=1e-12) for d in range(2))
            strong_ref = {"n_gpus": 1, "ms_per_step": ms1, "value": nq / (ms1 * 1e-3), "speedup": ms1 / (ms_dev_max / steps),
                          "efficiency": ms1 / (ms_dev_max / steps) / world, "results_equal_split_run": bool(same),
                          "what": "the same pair evaluated unsplit on rank 0's GPU (3 steps, L2 not flushed), for the strong-scaling ratio"}
        dist.barrier()

    # ---- end to end through the public API from pinned host float64 arrays
    e2e = None
    if not args.no_e2e and W.cfg == "2":
        def pinned(a):
            t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True)
            t.numpy()[...] = a
            return t.numpy()
        hA = Cloud(pinned(A.points), pinned(A.colors), pinned(A.normals))
        hB = Cloud(pinned(B.points), pinned(B.colors), pinned(B.normals))
        h2d = sum(x.nbytes for c in (hA, hB) for x in (c.points, c.colors, c.normals))
        n_e2e = max(3, steps // 2)
        ms_e2e, _, _, out = timed(lambda: step_e2e(hA, hB), n_e2e, 3, 0)
        ms_e2e = max_over_ranks(ms_e2e)
        _, _, tm3, _ = timed(lambda: step_e2e(hA, hB), 5, 1, 2)
        e2e = {"value": nq * n_e2e * per_rank_jobs / (ms_e2e * 1e-3), "unit": UNIT,
               "stage_ms_per_step": {k: round(v / 5, 5) for k, v in tm3.items() if k.endswith("_ms")},
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(2 * 96 + 16), "ms_per_step": ms_e2e / n_e2e,
               "api": "CloudPair(host float64 arrays) + MetricCalculator.calculate(transform_options(color=yuv, point_to_plane))"}
        # the same clouds in the compact form a voxelised PLY file holds (uint16 / uchar / float32): same API, informational
        def pinned_as(a, dt):
            t = torch.empty(a.shape, dtype={np.uint16: torch.uint16, np.uint8: torch.uint8, np.float32: torch.float32}[dt], pin_memory=True)
            o = t.numpy()
            o[...] = a.astype(dt)
            return o
        cA = Cloud(pinned_as(A.points, np.uint16), pinned_as(np.rint(A.colors * 255), np.uint8), pinned_as(A.normals, np.float32))
        cB = Cloud(pinned_as(B.points, np.uint16), pinned_as(np.rint(B.colors * 255), np.uint8), pinned_as(B.normals, np.float32))
        n_c = max(3, steps // 4)
        ms_c, _, _, out_c = timed(lambda: step_e2e(cA, cB), n_c, 3, 0)
        same = all(np.array_equal(np.asarray(out[k]), np.asarray(out_c[k])) for k in out if k[0] in ("GeoMSE", "ColorMSE") and k[-1] is not True)
        e2e["compact_inputs"] = {"value": nq * n_c / (ms_c * 1e-3), "ms_per_step": ms_c / n_c,
                                 "h2d_bytes_per_step": int(sum(x.nbytes for c in (cA, cB) for x in (c.points, c.colors, c.normals))),
                                 "what": "same API, clouds held as uint16 / uchar / float32 (float32 normals change D2 only)",
                                 "d1_and_colour_identical_to_float64_inputs": bool(same)}

    if not args.no_e2e and W.split:
        # the split pair through the same public API: EVERY rank is handed both clouds as pinned host float64 arrays
        # (that is the split's contract: a rank selects its slab on the device), uploads them inside the timed region,
        # evaluates its slab and takes part in the exchanges; the metrics come back on every rank
        try:
            def pinned_t(t):
                h = torch.empty(t.shape, dtype=torch.float64, pin_memory=True)
                h.copy_(t if t.dtype == torch.float64 else t.to(torch.float64) / 255.0)
                return h.numpy()
            hA = Cloud(pinned_t(dA[0]), pinned_t(dA[1]), pinned_t(dA[2]))
            hB = Cloud(pinned_t(dB[0]), pinned_t(dB[1]), pinned_t(dB[2]))
            torch.cuda.synchronize()
            h2d = sum(x.nbytes for c in (hA, hB) for x in (c.points, c.colors, c.normals))
            n_e2e = max(3, min(10, steps // 5))
            ms_e2e, _, _, out = timed(lambda: step_e2e(hA, hB), n_e2e, 3, 0)
            ms_e2e = max_over_ranks(ms_e2e)
            e2e = {"value": nq * n_e2e / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e / n_e2e,
                   "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(2 * 96 + 16),
                   "h2d_note": "per RANK: every rank uploads both clouds (the whole job moves this many bytes times the number of ranks)",
                   "api": "CloudPair(host float64 arrays, rank=, world=) + MetricCalculator.calculate(transform_options(color=yuv, point_to_plane)) "
                          "on every rank; partial records and boundary distances exchanged with NCCL inside the step"}
            del hA, hB
        except Exception as exc:       # (never lose the device-timed line to the informational arm)
            e2e = None
            print(f"[bench] e2e arm of the split pair failed on rank {rank}: {type(exc).__name__}: {exc}", file=sys.stderr, flush=True)

    if not args.no_e2e and W.cfg == "4":
        # configs[3]: the sequence API from pinned host float64 frames, pipelined (two contexts alternate: the uploads of
        # frame t+1 run under the kernels of frame t) against one frame at a time.  Several contexts = several streams: timed
        # with the wall clock between device-wide synchronisations.
        from open_pcc_metric_b200.sequence import evaluate_sequence

        def pinned(a):
            t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True)
            t.numpy()[...] = a
            return t.numpy()
        A2, B2 = W.gen(rank + 500)
        host = [(Cloud(pinned(A.points), pinned(A.colors), pinned(A.normals)), Cloud(pinned(B.points), pinned(B.colors), pinned(B.normals))),
                (Cloud(pinned(A2.points), pinned(A2.colors), pinned(A2.normals)), Cloud(pinned(B2.points), pinned(B2.colors), pinned(B2.normals)))]
        nframes = max(8, min(steps, 40))
        seq = [host[t % 2] for t in range(nframes)]
        qf = sum(len(a.points) + len(b.points) for a, b in seq)
        res_seq = {}
        for depth in (1, 2):
            evaluate_sequence(seq[:4], opts, ctx=ctx, pipeline=depth, peak="resolution", resolution_bits=W.bits)       # warm-up
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            evaluate_sequence(seq, opts, ctx=ctx, pipeline=depth, peak="resolution", resolution_bits=W.bits)
            torch.cuda.synchronize()
            dt = max_over_ranks((time.perf_counter() - t0) * 1e3)
            res_seq[depth] = dt
        h2d = sum(x.nbytes for a, b in seq for c in (a, b) for x in (c.points, c.colors, c.normals)) // nframes
        e2e = {"value": qf * per_rank_jobs / (res_seq[2] * 1e-3), "unit": UNIT, "ms_per_step": res_seq[2] / nframes,
               "frames_per_s": nframes * per_rank_jobs / (res_seq[2] * 1e-3), "frames": nframes * per_rank_jobs,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(2 * 96 + 16),
               "one_frame_at_a_time": {"ms_per_frame": res_seq[1] / nframes, "frames_per_s": nframes * per_rank_jobs / (res_seq[1] * 1e-3)},
               "api": "evaluate_sequence(frames of host float64 clouds, color=yuv, point_to_plane, pipeline=2): per-frame CloudPair + MetricCalculator, "
                      "frames round-robin over the ranks; wall clock between device-wide synchronisations"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not W.estimate_normals:
        from oracle import cpu_baseline as cb
        best = cb.cpu_best(A, B, W.color, True)
        cpu = {"value": best["queries_per_s"], "unit": UNIT, "cores": best["cores"], "kind": "port",
               "sample": "the FULL pair once: KD-tree builds + cKDTree.query(workers=-1) on all cores + whole-array numpy D1 / D2 / colour (measured, "
                         "not extrapolated); the reference's own single-threaded per-point loop structure is in reference_structure",
               "reference_structure": cpu_reference(W, A, B, args.cpu_sample)}

    if rank == 0:
        ps = psnr_summary(res, n_a, n_b, W.peak) if (W.peak and W.color) else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": args.warmup,
            "ms_per_step": ms_dev_max / steps, "higher_is_better": True,
            "scaling": "strong" if W.split else "weak", "vs_baseline": None,
            "dtype": "u32 distances / f64 epilogues" if W.bits else "f64 distances (float32 coordinates)",
            "data": "synthetic",
            "config": W.config(n_a, n_b, world, distinct_b),
            "ms_per_pair": ms_dev_max / steps,
            "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks,
            "gpu_launches": int(tm["total_launches"]), "library_launches": int(tm["library_launches"]),
            "wall_s_timed_region": wall_dev, "stage_ms_per_step": stages,
            "strong_scaling_reference": strong_ref,
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
