set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_vox.py -x -q > gpurun_out/vox_tests.log 2>&1; echo "vox tests rc=$?" 
tail -25 gpurun_out/vox_tests.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "gpu tests rc=$?"
tail -25 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_vox9.json 2> gpurun_out/bench_vox9.err; echo "bench rc=$?"
cat gpurun_out/bench_vox9.json; tail -5 gpurun_out/bench_vox9.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_brick_launches.csv python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/ncu_vox.log 2>&1
echo rc=$?
