"""GPU tier: the BENCHMARKED path at the BENCHMARKED sizes.

bench.py evaluates ``ctx.build_pair`` (occupancy-brick path) on
``synth_pair(10, 1_000_000, BASE_SEED + 2, dedup=False, oversample=4)``; the tests below run that very
call on that very pair and compare
  * the full per-point idx / d2 arrays of both directions with oracle tier O2
    (``o3d_standin.exact_knn``: cKDTree candidates, exact float64 recompute, smallest-index ties),
  * the fused reductions (sum / max of D1, D2, colour) with ``reference_port.PairOracle``
    (D1 bit-exact, D2 / colour rtol 1e-6 -- the tolerance BASELINE.json's north_star states),
and do the same for the shapes of configs[2] (vox12, >= 4 M, normals estimated) and configs[4]
(float32 LiDAR, >= 5 M) at sizes the oracle still finishes in tens of seconds.
Reference path being checked: cloud_pair.py:10-42, :54-80; metric.py:124-179, :213-228, :302-333, :353-366.
"""
import numpy as np
import pytest

from oracle import cnn, o3d_standin as o3s, reference_port as rp

pytestmark = pytest.mark.gpu

YUV = np.array([[0.25, 0.5, 0.25], [1, 0, -1], [-0.5, 1, -0.5]])


@pytest.fixture(scope="module")
def ctx():
    from open_pcc_metric_b200 import _native as N
    c = N.Context(0)
    yield c
    c.close()


def _tree_d2(search, query):
    """Exact squared NN distances for integer-valued clouds: cKDTree proposes the nearest point,
    the squared distance is recomputed from the coordinates (oracle tier O2, distances only)."""
    from scipy.spatial import cKDTree
    _, j = cKDTree(search, leafsize=15).query(query, k=1, workers=-1)
    d = query - search[j]
    return (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]


def test_bench_pair_brick_path_full_arrays_and_fused_sums(ctx):
    """configs[1] exactly as bench.py builds and evaluates it."""
    from open_pcc_metric_b200 import _native as N
    from open_pcc_metric_b200.synth import BASE_SEED, synth_pair
    A, B = synth_pair(10, 1_000_000, BASE_SEED + 2, step=2, dedup=False, oversample=4)
    assert len(A) == len(B)
    a = ctx.cloud(A.points, A.colors, A.normals)
    b = ctx.cloud(B.points, B.colors, B.normals)
    ctx.build_pair(a, b)                                     # the bench's call: brick index
    assert a.info().index_kind == 0
    ctx.reset_timings()
    res = ctx.pair_eval(a, b, N.EVAL_D2 | N.EVAL_COLOR | N.EVAL_PERPOINT, YUV, 1.0, N.NORMALS_BY_QUERY_INDEX)
    tm = ctx.timings()
    assert tm["vox_epilogue_ms"] >= 0 and tm["vox_tail"] > 400_000      # brick path ran; B's duplicate tail is in
    got = [(ctx.pair_get(N.GET_IDX, d, len(A)), ctx.pair_get(N.GET_D2, d, len(A))) for d in range(2)]
    fused_only = ctx.pair_eval(a, b, N.EVAL_D2 | N.EVAL_COLOR, YUV, 1.0, N.NORMALS_BY_QUERY_INDEX)   # what the bench times
    oracle = rp.PairOracle(A.points, B.points, A.colors, B.colors, A.normals, B.normals)
    for d, is_left in ((0, True), (1, False)):
        idx, d2 = got[d]
        assert np.array_equal(d2, oracle.d2[d]), d
        assert np.array_equal(idx, oracle.idx[d]), d
        for r in (res, fused_only):
            x = r.dir[d]
            assert x.n == x.n_total == len(A) and x.d1_exact_int == 1 and x.d2_valid == 1
            assert x.sum_d1_u64 == int(oracle.d2[d].sum()) and x.max_d1 == oracle.d2[d].max()
            pe2 = oracle.euclidean_distance(is_left, True)
            assert np.isclose(x.sum_d2, pe2.sum(), rtol=1e-6) and np.isclose(x.max_d2, pe2.max(), rtol=1e-6)
            cd2 = np.square(oracle.color_diff(is_left, "yuv"))
            assert np.allclose(list(x.color_sum), cd2.sum(0), rtol=1e-6)
            assert np.allclose(list(x.color_max), cd2.max(0), rtol=1e-6)
    assert bytes(res) == bytes(fused_only)
    # the de-duplicated variant of the same degradation: D1 + colour, full arrays
    _, Bd = synth_pair(10, 1_000_000, BASE_SEED + 2, step=2, dedup=True, oversample=4)
    bd = ctx.cloud(Bd.points, Bd.colors, Bd.normals)
    a2 = ctx.cloud(A.points, A.colors, A.normals)
    ctx.build_pair(a2, bd)
    r2 = ctx.pair_eval(a2, bd, N.EVAL_COLOR | N.EVAL_PERPOINT, YUV)
    o2 = rp.PairOracle(A.points, Bd.points, A.colors, Bd.colors, A.normals, Bd.normals)
    for d, is_left, n in ((0, True, len(A)), (1, False, len(Bd))):
        assert np.array_equal(ctx.pair_get(N.GET_D2, d, n), o2.d2[d])
        assert np.array_equal(ctx.pair_get(N.GET_IDX, d, n), o2.idx[d])
        assert r2.dir[d].sum_d1_u64 == int(o2.d2[d].sum())
        assert np.allclose(list(r2.dir[d].color_sum), np.square(o2.color_diff(is_left, "yuv")).sum(0), rtol=1e-6)
    for c in (a, b, a2, bd):
        c.close()


def test_config3_shape_vox12_4m_estimated_normals(ctx):
    """configs[2]: vox12, 4 M + 4 M points, no normals.  D1 / Hausdorff of the WHOLE pair are exact;
    neighbour indices and estimated normals (k = 30, Open3D recipe) on a 50 k sample against the C oracle."""
    from open_pcc_metric_b200 import _native as N
    from open_pcc_metric_b200.synth import BASE_SEED, degrade, synth_vox
    A = synth_vox(12, 4_000_000, BASE_SEED + 3, with_colors=False, with_normals=False, oversample=3)
    B = degrade(A, 2, BASE_SEED + 3, 12, dedup=False)
    a, b = ctx.cloud(A.points), ctx.cloud(B.points)
    ctx.build_pair(a, b)
    assert a.info().index_kind == 0
    res = ctx.pair_eval(a, b, N.EVAL_PERPOINT)
    rng = np.random.default_rng(3)
    for d, Q, S in ((0, A.points, B.points), (1, B.points, A.points)):
        want = _tree_d2(S, Q)
        d2 = ctx.pair_get(N.GET_D2, d, len(Q))
        assert np.array_equal(d2, want), d
        assert res.dir[d].sum_d1_u64 == int(want.sum()) and res.dir[d].max_d1 == want.max()      # GeoMSE / Hausdorff numerators
        sample = rng.choice(len(Q), 50_000, replace=False)
        oi, od = o3s.exact_knn(S, Q[sample], 1)
        assert np.array_equal(ctx.pair_get(N.GET_IDX, d, len(Q))[sample], oi[:, 0])
        assert np.array_equal(d2[sample], od[:, 0])
    # normals of cloud A, whole cloud estimated on the GPU, 50 k of them checked
    a.estimate_normals(30)
    got = a.get_normals()
    sample = rng.choice(len(A), 50_000, replace=False)
    ki, _ = o3s.exact_knn(A.points, A.points[sample], 30)
    want = cnn.normals(A.points, ki)
    dot = np.abs(np.sum(got[sample] * want, axis=1))
    assert (dot < 1 - 1e-9).mean() <= 0.002          # degenerate eigen-pairs have no defined direction
    assert np.allclose(np.linalg.norm(got[sample], axis=1), 1.0, atol=1e-12)
    a.close(); b.close()


def test_config5_shape_lidar_5m_float32(ctx):
    """configs[4] shape: float32 LiDAR-style pair, 5 M points; squared distances and indices of a
    100 k sample per direction are bit-exact against the float64 oracle (the reference holds float64
    copies of the float32 values)."""
    from open_pcc_metric_b200.synth import BASE_SEED, synth_lidar
    A, B = synth_lidar(5_000_000, BASE_SEED + 5)
    Ap, Bp = A.points.astype(np.float64), B.points.astype(np.float64)
    a, b = ctx.cloud(A.points), ctx.cloud(B.points)
    ctx.build_pair(a, b)
    assert a.info().index_kind == 1
    rng = np.random.default_rng(5)
    for q, s_, Q, S in ((a, b, Ap, Bp), (b, a, Bp, Ap)):
        idx, d2 = ctx.nn(q, s_)
        sample = rng.choice(len(Q), 100_000, replace=False)
        oi, od = o3s.exact_knn(S, Q[sample], 1)
        assert np.array_equal(d2[sample], od[:, 0])
        assert np.array_equal(idx[sample], oi[:, 0])
        # size-independent property: every reported d2 is reproduced by its own reported neighbour
        e = Q - S[idx]
        assert np.array_equal((e[:, 0] * e[:, 0] + e[:, 1] * e[:, 1]) + e[:, 2] * e[:, 2], d2)
    a.close(); b.close()
