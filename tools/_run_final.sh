set -x
cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_r1_final.json
timeout 300 python tools/run_configs.py c3 10000000 > gpurun_out/config3.log 2>&1; grep iter gpurun_out/config3.log | cut -c1-420
timeout 300 python tools/run_configs.py c5 > gpurun_out/config5.log 2>&1; grep iter gpurun_out/config5.log | cut -c1-420
timeout 300 python tools/run_configs.py c1 > gpurun_out/config1.log 2>&1; grep iter gpurun_out/config1.log | cut -c1-420
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_brick_launches.csv python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/ncu_vox.log 2>&1; echo rc=$?
