#!/usr/bin/env python
"""Range-checked run of the brick path (compute-sanitizer is not available on this pool): builds
libpccm with -DPCCM_VX_DEBUG -- every index the brick kernels compute is checked and violations are
counted on the device -- runs the GPU parity tests against that build and prints the counter.

    python tools/vx_debug_check.py          # on a GPU box; exit code 0 = tests green and 0 violations
"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "build", "libpccm_dbg.so")
os.makedirs(os.path.dirname(so), exist_ok=True)
subprocess.check_call([os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc"), "-DPCCM_VX_DEBUG", "-gencode", "arch=compute_100a,code=sm_100a",
                       "-O3", "-lineinfo", "--fmad=false", "-std=c++17", "-Xcompiler", "-fPIC,-O2", "-shared", "-o", so,
                       os.path.join(ROOT, "open_pcc_metric_b200", "csrc", "pccm_api.cu")])
os.environ["PCCM_LIB"] = so
sys.path.insert(0, ROOT)
os.chdir(ROOT)
import pytest  # noqa: E402

rc = int(pytest.main(["tests/test_gpu_vox.py", "tests/test_gpu_parity.py", "-x", "-q", "-m", "gpu", "-p", "no:cacheprovider"]))
from open_pcc_metric_b200 import _native as N  # noqa: E402

L = N.lib()
L.pccm_debug_errors.restype = ctypes.c_int
bad = L.pccm_debug_errors()
print(f"pytest rc {rc}; range-check violations counted by the debug build: {bad}")
sys.exit(1 if rc or bad else 0)
