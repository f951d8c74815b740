set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_vox.py -x -q > gpurun_out/vox_tests.log 2>&1; echo "vox tests rc=$?" 
tail -5 gpurun_out/vox_tests.log
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_vox5.json 2> gpurun_out/bench_vox5.err; echo "bench rc=$?"
cat gpurun_out/bench_vox5.json; tail -5 gpurun_out/bench_vox5.err
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "gpu tests rc=$?"
tail -5 gpurun_out/pytest_gpu.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_vox.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_vox.log 2>&1
echo rc=$?
ncu --set full --clock-control none --import-source on -k regex:"vx_search_kernel|vx_epilogue_kernel" -s 4 -c 2 -o gpurun_out/prof_vxq -f python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_vxq.log 2>&1
echo rc=$?
