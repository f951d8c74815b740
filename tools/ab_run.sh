# A/B of one switch on one box: GPU parity tests first, then bench lines with the switch off / on, alternating.
# usage: bash tools/ab_run.sh ENVVAR [extra bench args]
cd $GRAFT_REPO_ROOT
V=$1; shift
sum() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
r=d['roofline']
print('dev ms',round(d['ms_per_step'],4),'frac',round(r['frac'],4), 'stage ms', round(r.get('avg_launch_ms',0),4), {k:round(v,4) for k,v in r.get('launch_ms_by_kernel',{}).items()}, {k:round(v,4) for k,v in d['stage_ms_per_step'].items() if v}, d['gpu_launches'])"; }
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in 0 1 0 1; do
  echo "$V=$v"; env $V=$v python bench.py --steps 200 --no-cpu-baseline --no-e2e "$@" 2>/dev/null | sum
done
for sh in 0,8 3,8 7,8 1,4 0,2; do
  echo "split pair, rank,world = $sh"; python bench.py --config split --shard-of $sh --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | sum
done
