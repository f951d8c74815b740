cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_vox.py -x -q 2>&1 | tail -15
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -5
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
e=d['e2e']
print('ms_per_step',round(d['ms_per_step'],4), 'e2e', round(e['ms_per_step'],3), 'compact', round(e['compact_inputs']['ms_per_step'],3))
print({k:round(v,4) for k,v in e['stage_ms_per_step'].items() if v})"
