cd $GRAFT_REPO_ROOT
for v in 1 2; do
python bench.py --steps 200 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
e=d['e2e']
print('dev ms',round(d['ms_per_step'],4),'frac',round(d['roofline']['frac'],4), 'e2e', round(e['ms_per_step'],3), 'compact', round(e['compact_inputs']['ms_per_step'],3), {k:round(v,4) for k,v in d['stage_ms_per_step'].items() if v}, d['gpu_launches'])"
done
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
