"""CPU tier: the per-thread search logic of the CUDA kernels (pccm_core.cuh, compiled
for the host by tests/emul/emul.cpp) against the oracle -- every coordinate kind,
cell sizes from far too small to far too large, ties, duplicates, outliers."""
import ctypes
import os

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import cnn

_L = ctypes.CDLL(os.path.join(os.path.dirname(__file__), "emul", "libpccm_emul.so"))
KINT, KF32, KF64 = 0, 1, 2


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def emul_nn(kind, q, s, cell):
    q = np.ascontiguousarray(q, dtype=np.float64)
    s = np.ascontiguousarray(s, dtype=np.float64)
    idx = np.empty(len(q), np.int32)
    d2 = np.empty(len(q))
    _L.emul_nn(kind, _p(q), ctypes.c_int64(len(q)), _p(s), ctypes.c_int64(len(s)), ctypes.c_double(cell), _p(idx), _p(d2))
    return idx, d2


def emul_knn(kind, pts, k, cell):
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    n = len(pts)
    idx = np.empty((n, k), np.int32)
    d2 = np.empty((n, k))
    nr = np.empty((n, 3))
    _L.emul_knn_self(kind, _p(pts), ctypes.c_int64(n), k, ctypes.c_double(cell), _p(idx), _p(d2), _p(nr))
    return idx, d2, nr


def _check_nn(kind, q, s, cell):
    oi, od = cnn.knn(s, q, 1)
    i, d = emul_nn(kind, q, s, cell)
    assert np.array_equal(d, od[:, 0])
    assert np.array_equal(i, oi[:, 0])


@pytest.fixture(params=[0, 100000])
def short_row(request):
    """0 = binary search + sweeps (product default); huge = always the linear scan."""
    _L.emul_set_short_row(request.param)
    yield request.param
    _L.emul_set_short_row(0)


@pytest.mark.parametrize("cell", [1, 2, 4, 16, 256])
def test_int_nn_cells(cell, short_row):
    rng = np.random.default_rng(1)
    A = rng.integers(0, 64, (3000, 3)).astype(float)
    B = rng.integers(0, 64, (2000, 3)).astype(float)
    B[:5] += 300          # isolated far points: ring expansion must reach them
    B = np.concatenate([B, B[:50]])  # duplicates: smallest index must win
    _check_nn(KINT, A, B, cell)
    _check_nn(KINT, B, A, cell)


def test_int_extremes():
    A = np.array([[0, 0, 0], [32767, 32767, 32767], [0, 32767, 0], [16000, 1, 2]], dtype=float)
    B = np.array([[32767, 0, 32767], [1, 1, 1], [32766, 32767, 32767]], dtype=float)
    for cell in (1, 64, 4096, 32768):
        _check_nn(KINT, A, B, cell)
        _check_nn(KINT, B, A, cell)
    _check_nn(KINT, A, A[:1], 4)      # single-point search cloud
    _check_nn(KINT, A[:1], A, 4)


@pytest.mark.parametrize("kind", [KF32, KF64])
@pytest.mark.parametrize("cell", [0.01, 0.1, 0.5, 10.0])
def test_float_nn_cells(kind, cell, short_row):
    rng = np.random.default_rng(2)
    A = rng.normal(0, 1, (2500, 3))
    B = A[rng.permutation(2500)][:2000] + rng.normal(0, 0.05, (2000, 3))
    if kind == KF32:
        A = A.astype(np.float32).astype(float)
        B = B.astype(np.float32).astype(float)
    B[:3] += 40.0
    _check_nn(kind, A, B, cell)
    _check_nn(kind, B, A, cell)


def test_float_lattice_ties():
    """float kinds on integer-valued data: exact ties must still break on the index."""
    rng = np.random.default_rng(3)
    A = rng.integers(-20, 20, (1500, 3)).astype(float) * 0.5
    B = rng.integers(-20, 20, (1500, 3)).astype(float) * 0.5
    for kind in (KF32, KF64):
        for cell in (0.5, 2.0):
            _check_nn(kind, A, B, cell)


@pytest.mark.parametrize("kind,cell", [(KINT, 1), (KINT, 4), (KF32, 0.2), (KF64, 0.2)])
def test_knn_and_normals(kind, cell):
    rng = np.random.default_rng(4)
    if kind == KINT:
        P = rng.integers(0, 48, (2500, 3)).astype(float)
    else:
        P = rng.normal(0, 1, (2500, 3))
        if kind == KF32:
            P = P.astype(np.float32).astype(float)
    for k in (2, 30):
        oi, od = cnn.knn(P, P, k)
        i, d, nr = emul_knn(kind, P, k, cell)
        assert np.array_equal(d, od) and np.array_equal(i, oi)
        if k == 30:
            assert np.array_equal(nr, cnn.normals(P, oi))   # same arithmetic, same libm on the host


def test_knn_fewer_points_than_k():
    P = np.array([[0, 0, 0], [1, 0, 0], [5, 5, 5]], dtype=float)
    i, d, nr = emul_knn(KINT, P, 30, 2)
    assert np.array_equal(i[:, :3], [[0, 1, 2], [1, 0, 2], [2, 1, 0]]) and (i[:, 3:] == -1).all()
    assert np.isinf(d[:, 3:]).all()
    oi, _ = cnn.knn(P, P, 3)
    assert np.array_equal(nr, cnn.normals(P, oi))


coords = st.integers(min_value=0, max_value=40)
cloud = st.lists(st.tuples(coords, coords, coords), min_size=1, max_size=60)


@settings(max_examples=60, deadline=None, suppress_health_check=list(__import__('hypothesis').HealthCheck))
@given(a=cloud, b=cloud, shift=st.integers(0, 6))
def test_property_int_nn(a, b, shift):
    A = np.array(a, dtype=float)
    B = np.array(b, dtype=float)
    _check_nn(KINT, A, B, 1 << shift)


fl = st.floats(min_value=-50, max_value=50, allow_nan=False, width=32)
fcloud = st.lists(st.tuples(fl, fl, fl), min_size=1, max_size=40)


@settings(max_examples=40, deadline=None, suppress_health_check=list(__import__('hypothesis').HealthCheck))
@given(a=fcloud, b=fcloud, cell=st.sampled_from([0.05, 1.0, 30.0]), kind=st.sampled_from([KF32, KF64]))
def test_property_float_nn(a, b, cell, kind):
    _check_nn(kind, np.array(a, dtype=float), np.array(b, dtype=float), cell)


# ---- occupancy-brick path (pccm_vox.cuh): build functions + staged / general search ----------
def emul_vox_nn(q, s, max_ring=2):
    q = np.ascontiguousarray(q, dtype=np.float64)
    s = np.ascontiguousarray(s, dtype=np.float64)
    idx = np.empty(len(q), np.int32)
    d2 = np.empty(len(q))
    stats = np.zeros(5, np.int64)
    rc = _L.emul_vox_nn(_p(q), ctypes.c_int64(len(q)), _p(s), ctypes.c_int64(len(s)), ctypes.c_int(max_ring), _p(idx), _p(d2), _p(stats))
    assert rc in (0, 1), f"internal consistency check {rc} failed"
    return (idx, d2, stats) if rc == 0 else (None, None, None)


def _check_vox(q, s, max_ring=2):
    oi, od = cnn.knn(s, q, 1)
    i, d, stats = emul_vox_nn(q, s, max_ring)
    assert stats is not None, "brick grid over the directory budget"
    assert np.array_equal(d, od[:, 0])
    assert np.array_equal(i, oi[:, 0])
    return stats


def test_vox_dense_ties_duplicates_outliers():
    rng = np.random.default_rng(1)
    A = rng.integers(0, 64, (3000, 3)).astype(float)
    B = rng.integers(0, 64, (2000, 3)).astype(float)
    B[:5] += 300                      # isolated far points: the general search must reach them
    B = np.concatenate([B, B[:50]])   # duplicates: smallest index wins, the others form the tail
    sa = _check_vox(A, B)
    sb = _check_vox(B, A)
    assert sa[0] > 0 and sa[1] > 0 and sa[2] > 0      # every stage of the search is exercised
    assert sb[3] >= 50                                 # the duplicate tail is queried too


def test_vox_lattice_pair_is_decided_by_the_staged_rows():
    rng = np.random.default_rng(5)
    A = np.unique(rng.integers(100, 140, (4000, 3)), axis=0).astype(float)
    J = rng.integers(-1, 2, A.shape) * (rng.random(A.shape) < 0.25)
    B = np.unique((np.round(A / 2) * 2 + J).clip(0, 1023), axis=0)
    rng.shuffle(A)
    rng.shuffle(B)
    for q, s in ((A, B), (B, A)):
        stats = _check_vox(q, s)
        assert stats[3] == 0 and stats[2] < 0.05 * len(q)     # (rounding + jitter can move a point 2 voxels per axis)


def test_vox_extremes_and_far_apart():
    A = np.array([[0, 0, 0], [4095, 4095, 4095], [0, 4095, 0], [2000, 1, 2]], dtype=float)
    B = np.array([[4095, 0, 4095], [1, 1, 1], [4094, 4095, 4095]], dtype=float)
    for ring in (0, 2, 1000):                 # 1000: the brick rings finish every query themselves
        _check_vox(A, B, ring)
        _check_vox(B, A, ring)
    _check_vox(A, A[:1])
    _check_vox(A[:1], A)
    full = np.array([[0, 0, 0], [32767, 32767, 32767]], dtype=float)
    assert emul_vox_nn(full, full)[2] is None  # 15-bit cube: over the directory budget -> pencil path only
    thin = np.array([[0, 0, 0], [32767, 900, 900], [31000, 5, 7]], dtype=float)
    _check_vox(thin, thin[::-1].copy())        # one long axis fits
    rng = np.random.default_rng(6)
    P = rng.integers(0, 200, (2000, 3)).astype(float)
    s = _check_vox(P, P + [0, 250, 0])         # disjoint clouds: everything ends in the pencil search
    assert s[4] > 0
    s = _check_vox(P, P + [0, 250, 0], 1000)
    assert s[4] == 0
    _check_vox(P, P[::-1].copy())              # identical clouds: d2 = 0 everywhere
    s = _check_vox(np.tile([[3.0, 3, 3]], (100, 1)), np.tile([[3.0, 4, 3]], (40, 1)))   # one voxel each, long groups
    assert s[3] == 99
    W = rng.integers(0, 4096, (2000, 3)).astype(float)
    _check_vox(W, (W + rng.integers(-3, 4, W.shape)).clip(0, 4095))
    _check_vox(W, rng.integers(0, 4096, (500, 3)).astype(float))   # sparse: rings, then pencils


def test_vox_window_edges():
    """neighbours exactly 15 / 16 / 17 voxels away along x and across brick borders in y, z"""
    base = np.array([[31, 7, 7], [32, 8, 8], [0, 0, 0], [63, 15, 16]], dtype=float)
    for dx in (-17, -16, -15, -2, -1, 1, 2, 15, 16, 17):
        for dy, dz in ((0, 0), (1, 0), (0, -1), (2, 2), (-2, 1), (3, 0)):
            S = (base + 40 + [dx, dy, dz])
            Q = base + 40
            _check_vox(Q, S)
            _check_vox(S, Q)


@settings(max_examples=80, deadline=None, suppress_health_check=list(__import__('hypothesis').HealthCheck))
@given(a=cloud, b=cloud, off=st.tuples(st.integers(0, 3000), st.integers(0, 3000), st.integers(0, 3000)))
def test_property_vox_nn(a, b, off):
    A = np.array(a, dtype=float) + np.array(off, dtype=float)
    B = np.array(b, dtype=float) + np.array(off, dtype=float)
    _check_vox(A, B)


def test_vox_boundary_distances():
    """nearest OTHER point by the staged row scans with the own bit cleared (vx_selfnn_kernel's logic)"""
    _L.emul_vox_self.restype = ctypes.c_int64
    rng = np.random.default_rng(9)
    surf = np.unique(rng.integers(0, 48, (6000, 3)), axis=0).astype(float)        # dense: everything decided by the rows
    sparse = rng.integers(0, 600, (400, 3)).astype(float)                          # mostly isolated points
    edges = np.array([[31, 7, 7], [32, 7, 7], [47, 8, 8], [0, 0, 0], [16, 0, 0], [17, 0, 0], [63, 15, 16], [63, 15, 18]], dtype=float)
    for P in (surf, np.concatenate([surf, surf[:200]]), sparse, edges):
        P = np.ascontiguousarray(P)
        out = np.empty(len(P))
        und = _L.emul_vox_self(_p(P), ctypes.c_int64(len(P)), _p(out))
        assert und >= 0
        _, o2 = cnn.knn(P, P, 2)
        decided = out >= 0
        assert np.array_equal(out[decided], o2[decided, 1])
        assert (o2[~decided, 1] >= 9).all()              # only answers 3+ voxels away are left to the brick scans
        assert und == (~decided).sum()
    assert decided.sum() >= 5
