cd $GRAFT_REPO_ROOT
set -x
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_r1b.json 2> gpurun_out/bench_r1b.err; echo "bench rc=$?"; cat gpurun_out/bench_r1b.json
timeout 300 python tools/run_configs.py c3 10000000 > gpurun_out/config3.log 2>&1; tail -4 gpurun_out/config3.log | cut -c1-400
timeout 300 python tools/run_configs.py c1 > gpurun_out/config1.log 2>&1; tail -3 gpurun_out/config1.log | cut -c1-300
timeout 300 python tools/far_apart.py > gpurun_out/far_apart.log 2>&1; cat gpurun_out/far_apart.log
