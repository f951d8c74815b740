cd $GRAFT_REPO_ROOT
for v in "" build/libpccm_noink.so; do
echo "== $v"
PCCM_LIB=${v:+$PWD/$v} 
if [ -n "$v" ]; then export PCCM_LIB=$PWD/$v; else unset PCCM_LIB; fi
timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms_per_step',round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['stage_ms_per_step'].items() if v}, d['brick_path'])"
done
