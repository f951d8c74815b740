#!/usr/bin/env python
"""Per-warp timeline of vx_search_kernel on the bench pair (library built with -DPCCM_VX_TRACE: tools/variants.py build
trace:-DPCCM_VX_TRACE): when do the warps finish, how many bricks / voxels did each take?"""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("PCCM_LIB", os.path.join(ROOT, "build", "libpccm_trace.so"))
from open_pcc_metric_b200 import _native as N, synth  # noqa: E402
A, B = synth.synth_pair(10, 1_000_000, synth.BASE_SEED + 2, step=2, dedup=False, oversample=4)
dev = torch.device("cuda:0")
dA = [torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (A.points, (A.colors * 255).round().astype(np.uint8), A.normals)]
dB = [torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (B.points, (B.colors * 255).round().astype(np.uint8), B.normals)]
ctx = N.Context(0)
YUV = np.array([[0.25, 0.5, 0.25], [1, 0, -1], [-0.5, 1, -0.5]])
for it in range(4):
    a = ctx.cloud(*dA); b = ctx.cloud(*dB); ctx.build_pair(a, b)
    ctx.pair_eval(a, b, N.EVAL_D2 | N.EVAL_COLOR, YUV); a.close(); b.close()
L = N.lib()
nw = 148 * 10 * 4
buf = np.zeros((nw, 4), dtype=np.uint64)
L.pccm_debug_trace.restype = ctypes.c_int
got = L.pccm_debug_trace(buf.ctypes.data_as(ctypes.c_void_p), nw)
assert got == nw, got
t0 = buf[:, 0].min()
start = (buf[:, 0] - t0).astype(np.float64) / 1e3
end = (buf[:, 1] - t0).astype(np.float64) / 1e3
print(f"warps {nw}: start median {np.median(start):.1f} max {start.max():.1f} us; end min {end.min():.1f} p10 {np.percentile(end, 10):.1f} "
      f"median {np.median(end):.1f} p90 {np.percentile(end, 90):.1f} max {end.max():.1f} us")
print("mean active fraction of the kernel's duration:", float(((end - start).sum()) / (nw * end.max())))
br, vx = buf[:, 2].astype(np.int64), buf[:, 3].astype(np.int64)
print(f"bricks per warp: min {br.min()} median {int(np.median(br))} max {br.max()}; voxels per warp: min {vx.min()} median {int(np.median(vx))} max {vx.max()}")
late = np.argsort(-end)[:8]
print("last warps: " + ", ".join(f"end {end[i]:.1f} us bricks {br[i]} voxels {vx[i]}" for i in late))
hist, edges = np.histogram(end, bins=12)
print("finish-time histogram:", [(round(float(e), 1), int(h)) for e, h in zip(edges, hist)])
