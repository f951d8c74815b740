cd $GRAFT_REPO_ROOT
timeout 300 python tools/run_configs.py c3 10000000 2>&1 | grep iter | cut -c1-200
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "normals or metrics or cli" 2>&1 | tail -3
