cd $GRAFT_REPO_ROOT
for v in base e5 e4 e6g2 e4p4 e5g2; do
if [ $v = base ]; then L=$PWD/open_pcc_metric_b200/libpccm.so; else L=$PWD/build/libpccm_$v.so; fi
PCCM_LIB=$L python bench.py --steps 200 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v dev ms',round(d['ms_per_step'],4),'frac',round(d['roofline']['frac'],4), {k:round(v,4) for k,v in d['stage_ms_per_step'].items() if v}, d['check'])"
done
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/e2e_timeline.py 2>&1 | grep -A3 "300 eval"
