import os, sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from open_pcc_metric_b200 import _native as N, synth
A, B = synth.synth_pair(10, 1_000_000, synth.BASE_SEED + 2, step=2, dedup=False, oversample=4)
def pinned(a):
    t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True); t.numpy()[...] = a; return t
hp = pinned(A.points)
d = torch.empty_like(hp, device='cuda')
torch.cuda.synchronize()
for _ in range(3):
    t0 = time.perf_counter(); d.copy_(hp, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print('torch pinned H2D 24MB: %.3f ms = %.1f GB/s' % (dt * 1e3, hp.numel() * 8 / dt / 1e9))
ctx = N.Context(0)
for label, arr in (('pinned f64', hp.numpy()), ('pageable f64', A.points), ('pinned u16', None)):
    if arr is None:
        t16 = torch.empty(A.points.shape, dtype=torch.uint16, pin_memory=True); t16.numpy()[...] = A.points.astype(np.uint16); arr = t16.numpy()
    for rep in range(4):
        ctx.synchronize(); t0 = time.perf_counter()
        c = ctx.cloud(arr); t1 = time.perf_counter(); ctx.synchronize(); t2 = time.perf_counter()
        c.close()
    print(label, 'cloud(): call %.3f ms, until done %.3f ms' % ((t1 - t0) * 1e3, (t2 - t0) * 1e3))
