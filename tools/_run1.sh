set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_vox.py -x -q > gpurun_out/vox_tests.log 2>&1; echo "vox tests rc=$?" 
tail -30 gpurun_out/vox_tests.log
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_vox3.json 2> gpurun_out/bench_vox3.err; echo "bench rc=$?"
cat gpurun_out/bench_vox3.json; tail -5 gpurun_out/bench_vox3.err
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "gpu tests rc=$?"
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python tools/far_apart.py > gpurun_out/far_apart.log 2>&1; echo "far rc=$?"; cat gpurun_out/far_apart.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_vox.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_vox.log 2>&1
echo rc=$?
