# A/B on one box: voxel coordinates from the occupancy rows (PCCM_VXYZ_ROWS), compacting epilogue with two gathers per thread
cd $GRAFT_REPO_ROOT
sum() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
r=d['roofline']
print('dev ms',round(d['ms_per_step'],4),'frac',round(r['frac'],4), 'stage ms', round(r.get('avg_launch_ms',0),4), {k:round(v,4) for k,v in d['stage_ms_per_step'].items() if v}, d['gpu_launches'])"; }
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
B="python bench.py --no-cpu-baseline --no-e2e"
for v in 0 1; do echo "1M pair PCCM_VXYZ_ROWS=$v"; PCCM_VXYZ_ROWS=$v $B --steps 200 2>/dev/null | sum; done
S="$B --config split --steps 10 --warmup 3"
for v in 0 1; do echo "10M whole PCCM_VXYZ_ROWS=$v"; PCCM_VXYZ_ROWS=$v $S 2>/dev/null | sum; done
for v in 0 1; do echo "3,8 PCCM_VXYZ_ROWS=$v"; PCCM_VXYZ_ROWS=$v $S --shard-of 3,8 2>/dev/null | sum; done
