cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py 2>&1 | grep -v Warning | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 100 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_n2.json; python -c "
import json; d=json.load(open('gpurun_out/bench_n2.json')); print('N=2 value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'] if d.get('e2e') else None)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --mode partition --steps 100 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 > gpurun_out/bench_n2_partition.json; python -c "
import json; d=json.load(open('gpurun_out/bench_n2_partition.json')); print('N=2 partition value',d['value'],'ms',d['ms_per_step'])"
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N=1 value',d['value'],'ms',d['ms_per_step'], d['stage_ms_per_step'])"
