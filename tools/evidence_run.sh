cd $GRAFT_REPO_ROOT
python tools/e2e_bisect.py 2>&1 | tail -8
PCCM_BENCH_VERBOSE=1 python bench.py --steps 200 --no-cpu-baseline 2>gpurun_out/b.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
e=d['e2e']
print('dev ms',round(d['ms_per_step'],4),'frac',round(d['roofline']['frac'],4),'e2e ms',round(e['ms_per_step'],3), 'compact ms', round(e['compact_inputs']['ms_per_step'],3), {k:round(v,3) for k,v in e['stage_ms_per_step'].items() if v})
print({k:round(v,4) for k,v in d['stage_ms_per_step'].items() if v}, d['gpu_launches'])"
grep timed gpurun_out/b.err
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
