"""ctypes loader for the plain-C oracle (oracle/pccm_oracle.c) -- TEST INFRASTRUCTURE."""
from __future__ import annotations

import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libpccm_oracle.so")
    src = os.path.join(_HERE, "pccm_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libpccm_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.oracle_knn.restype = ctypes.c_int
        _LIB.oracle_normals.restype = ctypes.c_int
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def num_threads():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def _fan(n, threads, fn):
    threads = threads or num_threads()
    if threads <= 1 or n < 256:
        assert fn(0, n) == 0
        return
    step = max(64, -(-n // (threads * 8)))
    ranges = [(b, min(n, b + step)) for b in range(0, n, step)]
    with ThreadPoolExecutor(threads) as ex:
        for rc in ex.map(lambda r: fn(*r), ranges):
            assert rc == 0


def knn(points, queries, k, threads=None):
    """Canonical exact k-NN (rows sorted by (d2, index)).  Returns idx[Q,k], d2[Q,k]."""
    pts = np.ascontiguousarray(points, dtype=np.float64)
    q = np.ascontiguousarray(queries, dtype=np.float64)
    nq = len(q)
    idx = np.empty((nq, k), dtype=np.int64)
    d2 = np.empty((nq, k), dtype=np.float64)
    L = lib()
    _fan(nq, threads, lambda b, e: L.oracle_knn(_p(pts), ctypes.c_int64(len(pts)), _p(q), ctypes.c_int64(b),
                                                ctypes.c_int64(e), ctypes.c_int(k), _p(idx), _p(d2)))
    return idx, d2


def normals(points, nn_idx, threads=None):
    """Normal of every ROW of nn_idx (neighbour lists into ``points``; one row per point of the cloud,
    or the rows of a sample of it)."""
    pts = np.ascontiguousarray(points, dtype=np.float64)
    nn = np.ascontiguousarray(nn_idx, dtype=np.int64)
    out = np.empty((len(nn), 3), dtype=np.float64)
    L = lib()
    _fan(len(nn), threads, lambda b, e: L.oracle_normals(_p(pts), ctypes.c_int64(b), ctypes.c_int64(e), _p(nn),
                                                          ctypes.c_int(nn.shape[1]), _p(out)))
    return out
