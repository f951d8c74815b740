cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_vox.py -x -q 2>&1 | tail -2
timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms_per_step',round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['stage_ms_per_step'].items() if v}, d['brick_path'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_brick_launches.csv python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/ncu_vox.log 2>&1
