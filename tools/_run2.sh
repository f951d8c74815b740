cd $GRAFT_REPO_ROOT
python - <<'PY'
import cProfile, pstats, sys, time, io
sys.path.insert(0, '.')
import numpy as np
from open_pcc_metric_b200 import _native as N, synth
from open_pcc_metric_b200.calculator import MetricCalculator
from open_pcc_metric_b200.cloud_pair import CloudPair
from open_pcc_metric_b200.options import CalculateOptions, transform_options
ctx = N.Context(0)
A = synth.synth_vox(12, 10_000_000, synth.BASE_SEED + 3, with_colors=False, with_normals=False, oversample=3)
B = synth.degrade(A, 2, synth.BASE_SEED + 3, 12, dedup=False)
opts = CalculateOptions(color=None, hausdorff=True, point_to_plane=True)
def run():
    p = CloudPair(A, B, ctx=ctx, peak="resolution", resolution_bits=12)
    r = MetricCalculator(p).calculate(transform_options(opts)).as_dict()
    p.close()
    return r
run()
A.normals = None; B.normals = None
t = time.perf_counter(); 
pr = cProfile.Profile(); pr.enable(); run(); pr.disable()
print("total ms", (time.perf_counter() - t) * 1e3)
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(14); print(s.getvalue()[:3500])
PY
