set -x
cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err; echo "bench rc=$?"; cat gpurun_out/bench_r1_final.json
PCCM_VOX=0 timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_r1_pencil.json 2>/dev/null; echo "pencil bench rc=$?"; cat gpurun_out/bench_r1_pencil.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_brick_launches.csv python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/ncu_vox.log 2>&1; echo rc=$?
ncu --set full --clock-control none --import-source on -k regex:"vx_search_kernel|vx_epilogue_kernel|vx_general_kernel" -s 6 -c 3 -o gpurun_out/prof_brick -f python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/ncu_vxq.log 2>&1; echo rc=$?
timeout 300 python tools/run_configs.py c3 10000000 > gpurun_out/config3.log 2>&1; tail -14 gpurun_out/config3.log | cut -c1-420
timeout 300 python tools/run_configs.py c1 > gpurun_out/config1.log 2>&1; tail -9 gpurun_out/config1.log | cut -c1-300
timeout 300 python tools/far_apart.py > gpurun_out/far_apart.log 2>&1; cat gpurun_out/far_apart.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1_ref.json 2>/dev/null; echo "ref rc=$?"; cat gpurun_out/bench_r1_ref.json | cut -c1-600
