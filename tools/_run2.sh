cd $GRAFT_REPO_ROOT
for v in mb8 mb10 mb12; do
echo "== $v"
PCCM_LIB=$PWD/build/libpccm_$v.so timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms_per_step',round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['stage_ms_per_step'].items() if v}, d['roofline']['launch_ms_by_kernel'])"
done
