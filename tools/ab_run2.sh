# A/B on one box: bulk-copy statistics pass, compacting epilogue of split pairs; launch list of the 10 M pair
cd $GRAFT_REPO_ROOT
sum() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
r=d['roofline']
print('dev ms',round(d['ms_per_step'],4),'frac',round(r['frac'],4), 'stage ms', round(r.get('avg_launch_ms',0),4), {k:round(v,4) for k,v in d['stage_ms_per_step'].items() if v}, d['gpu_launches'])"; }
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
B="python bench.py --no-cpu-baseline --no-e2e"
for v in 0 1; do echo "1M pair PCCM_STATS_TMA=$v"; PCCM_STATS_TMA=$v $B --steps 200 2>/dev/null | sum; done
S="$B --config split --steps 10 --warmup 3"
echo "3,8 compact=0 stats_tma=0"; PCCM_EPI_COMPACT=0 PCCM_STATS_TMA=0 $S --shard-of 3,8 2>/dev/null | sum
echo "3,8 compact=1 stats_tma=0"; PCCM_EPI_COMPACT=1 PCCM_STATS_TMA=0 $S --shard-of 3,8 2>/dev/null | sum
echo "3,8 compact=1 stats_tma=1"; $S --shard-of 3,8 2>/dev/null | sum
echo "0,8"; $S --shard-of 0,8 2>/dev/null | sum
echo "1,4"; $S --shard-of 1,4 2>/dev/null | sum
echo "0,2"; $S --shard-of 0,2 2>/dev/null | sum
echo "whole"; $S 2>/dev/null | sum
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_10m.csv python bench.py --config split --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_l10m.log 2>&1; echo "ncu list rc=$?"
