"""GPU tier: the occupancy-brick path (pccm_vox*.cuh) of integer pairs against the brute-force C
oracle and against the pencil path of the same library (PCCM_VOX=0) -- staged bit-scan search,
general search, pencil fallback for far queries, duplicate tails, rank slices, determinism."""
import os

import numpy as np
import pytest

from oracle import cnn

pytestmark = pytest.mark.gpu

YUV = np.array([[0.25, 0.5, 0.25], [1, 0, -1], [-0.5, 1, -0.5]])


@pytest.fixture(scope="module")
def ctxs():
    from open_pcc_metric_b200 import _native as N
    vox = N.Context(0)
    os.environ["PCCM_VOX"] = "0"
    try:
        pencil = N.Context(0)
    finally:
        del os.environ["PCCM_VOX"]
    yield vox, pencil
    vox.close()
    pencil.close()


def _attrs(rng, n):
    col = rng.integers(0, 256, (n, 3)).astype(np.float64) / 255.0
    nrm = rng.normal(0, 1, (n, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    return col, nrm


def _cases():
    from open_pcc_metric_b200.synth import synth_pair
    rng = np.random.default_rng(77)
    out = {}
    A, B = synth_pair(9, 60000, 11)
    out["surface"] = (A.points, B.points)
    P = rng.integers(0, 64, (5000, 3)).astype(np.float64)
    Q = rng.integers(0, 64, (4000, 3)).astype(np.float64)
    Q[:7] += 400
    out["dense_dups_outliers"] = (np.concatenate([P, P[:300], P[:100]]), np.concatenate([Q, Q[-200:]]))
    S = rng.integers(0, 300, (6000, 3)).astype(np.float64)
    out["far_apart"] = (S, S + [0, 330, 40])
    out["sparse"] = (rng.integers(0, 4096, (3000, 3)).astype(np.float64), rng.integers(0, 4096, (2500, 3)).astype(np.float64))
    L = np.unique(rng.integers(200, 260, (20000, 3)), axis=0).astype(np.float64)
    out["lattice"] = (L, np.unique(np.round(L / 4) * 4, axis=0))
    out["one_point_each"] = (np.array([[5.0, 6, 7]]), np.array([[900.0, 2, 40]]))
    out["same_voxel_only"] = (np.tile([[3.0, 3, 3]], (700, 1)), np.tile([[3.0, 4, 3]], (40, 1)))
    return out


CASES = None


def _case(name):
    global CASES
    if CASES is None:
        CASES = _cases()
    return CASES[name]


NAMES = ["surface", "dense_dups_outliers", "far_apart", "sparse", "lattice", "one_point_each", "same_voxel_only"]


@pytest.mark.parametrize("name", NAMES)
def test_vox_nn_equals_oracle(ctxs, name):
    vox, _ = ctxs
    A, B = _case(name)
    a, b = vox.cloud(A), vox.cloud(B)
    vox.build_pair(a, b)
    for q, s, Q, S in ((a, b, A, B), (b, a, B, A)):
        idx, d2 = vox.nn(q, s)
        oi, od = cnn.knn(S, Q, 1)
        assert np.array_equal(d2, od[:, 0]), name
        assert np.array_equal(idx, oi[:, 0]), name
    a.close(); b.close()


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("mode", [0, 1])
def test_vox_pair_eval_equals_pencil_path(ctxs, name, mode):
    """every fused quantity and the per-point products, both normal modes"""
    from open_pcc_metric_b200 import _native as N
    A, B = _case(name)
    rng = np.random.default_rng(5)
    n = min(len(A), len(B))          # reference normal indexing needs equal lengths (quirk Q1)
    A, B = A[:n], B[:n]
    ca, na = _attrs(rng, n)
    cb, nb = _attrs(rng, n)
    flags = N.EVAL_D2 | N.EVAL_COLOR | N.EVAL_PERPOINT
    res = []
    for ctx in ctxs:
        a, b = ctx.cloud(A, ca, na), ctx.cloud(B, cb, nb)
        ctx.build_pair(a, b)
        r = ctx.pair_eval(a, b, flags, YUV, 1.0, mode)
        pp = [(ctx.pair_get(N.GET_IDX, d, n), ctx.pair_get(N.GET_D2, d, n)) for d in range(2)]
        res.append((r, pp, ctx.timings()))
        a.close(); b.close()
    (rv, pv, tv), (rp, pq, _) = res
    for d in range(2):
        assert np.array_equal(pv[d][0], pq[d][0]) and np.array_equal(pv[d][1], pq[d][1]), (name, d)
        x, y = rv.dir[d], rp.dir[d]
        assert (x.n, x.n_total, x.sum_d1_u64, x.max_d1, x.max_d2, x.d2_valid) == (y.n, y.n_total, y.sum_d1_u64, y.max_d1, y.max_d2, y.d2_valid)
        assert x.sum_d1_u64 == int(pv[d][1].sum())
        assert np.isclose(x.sum_d2, y.sum_d2, rtol=1e-12, atol=0)
        for c in range(3):
            assert np.isclose(x.color_sum[c], y.color_sum[c], rtol=1e-12, atol=0)
            assert x.color_max[c] == y.color_max[c]
    if name == "far_apart":
        assert tv["vox_far"] > 0
    if name == "dense_dups_outliers":
        assert tv["vox_tail"] > 0 and tv["vox_undecided"] > 0


def test_vox_rank_slices_add_up_and_are_deterministic(ctxs):
    from open_pcc_metric_b200 import _native as N
    vox, _ = ctxs
    A, B = _case("dense_dups_outliers")
    n = min(len(A), len(B))
    A, B = A[:n], B[:n]
    rng = np.random.default_rng(6)
    ca, na = _attrs(rng, n)
    cb, nb = _attrs(rng, n)
    outs = []
    for rep in range(3):
        a, b = vox.cloud(A, ca, na), vox.cloud(B, cb, nb)
        vox.build_pair(a, b)
        full = vox.pair_eval(a, b, N.EVAL_D2 | N.EVAL_COLOR, YUV)
        outs.append(bytes(full))
        if rep == 0:
            for world in (2, 3, 7):
                parts = [vox.pair_eval(a, b, N.EVAL_D2 | N.EVAL_COLOR, YUV, rank=r, world=world) for r in range(world)]
                for d in range(2):
                    assert sum(p.dir[d].n for p in parts) == full.dir[d].n == n
                    assert sum(p.dir[d].sum_d1_u64 for p in parts) == full.dir[d].sum_d1_u64
                    assert max(p.dir[d].max_d1 for p in parts) == full.dir[d].max_d1
                    assert max(p.dir[d].max_d2 for p in parts) == full.dir[d].max_d2
                    assert np.isclose(sum(p.dir[d].sum_d2 for p in parts), full.dir[d].sum_d2, rtol=1e-12)
                    for c in range(3):
                        assert np.isclose(sum(p.dir[d].color_sum[c] for p in parts), full.dir[d].color_sum[c], rtol=1e-12)
        a.close(); b.close()
    assert outs[0] == outs[1] == outs[2]     # duplicate tails and undecided queries included


@pytest.mark.parametrize("name", NAMES)
def test_vox_boundary_distances_equal_oracle_and_pencil_path(ctxs, name):
    """compute_nearest_neighbor_distance (cloud_pair.py:108-109) on the brick index: the pair search's
    row scans with the voxel's own bit cleared; isolated points make the call fall back to the pencil path"""
    vox, pencil = ctxs
    A, B = _case(name)
    if len(A) < 2:
        pytest.skip("needs two points")
    res = []
    for ctx in (vox, pencil):
        a, b = ctx.cloud(A), ctx.cloud(B)
        ctx.build_pair(a, b)
        mn, mx, per = a.self_nn_minmax(per_point=True)
        parts = [a.self_nn_minmax(len(A) * r // 3, len(A) * (r + 1) // 3)[:2] for r in range(3)]
        res.append((mn, mx, per, parts))
        a.close(); b.close()
    _, o2 = cnn.knn(A, A, 2)
    want = np.sqrt(o2[:, 1])
    for mn, mx, per, parts in res:
        assert np.array_equal(per, want), name
        assert mn == want.min() and mx == want.max()
        assert min(p[0] for p in parts) == mn and max(p[1] for p in parts) == mx


def test_attach_equals_create_with_attributes(ctxs):
    """pccm_cloud_attach (coordinates of both clouds first, colours / normals afterwards on the copy
    stream) gives the same evaluation as pccm_cloud_create with everything; float colours that are
    not k/255 take the float64 colour array."""
    from open_pcc_metric_b200 import _native as N
    vox, _ = ctxs
    A, B = _case("surface")
    n = min(len(A), len(B))
    A, B = A[:n], B[:n]
    rng = np.random.default_rng(8)
    for exact_u8 in (True, False):
        ca, na = _attrs(rng, n)
        cb, nb = _attrs(rng, n)
        if not exact_u8:
            ca, cb = rng.random((n, 3)), rng.random((n, 3))
        outs = []
        for late in (0, 1, 2):
            if late == 0:
                a, b = vox.cloud(A, ca, na), vox.cloud(B, cb, nb)
                vox.build_pair(a, b)
            elif late == 1:
                a, b = vox.cloud(A), vox.cloud(B)
                a.attach(ca, na); b.attach(cb, nb)
                vox.build_pair(a, b)
            else:                                   # a brick-indexed pair accepts colours until its first colour evaluation
                a, b = vox.cloud(A), vox.cloud(B)
                vox.build_pair(a, b)
                a.attach(ca, na); b.attach(cb, nb)
            assert bool(a.info().colors_u8) == exact_u8
            outs.append(bytes(vox.pair_eval(a, b, N.EVAL_D2 | N.EVAL_COLOR, YUV)))
            a.close(); b.close()
        assert outs[0] == outs[1] == outs[2]
    a = vox.cloud(A)
    a.build_index()
    with pytest.raises(N.PccmError):
        a.attach(ca, None)          # pencil records bake the colours in at build time
    a.close()


def test_brick_indexed_cloud_outlives_its_partner_and_joins_other_pairs(ctxs):
    """the pencil index of a brick-indexed cloud is built on demand from the pair's records: also after
    the partner is gone, and for evaluations against a cloud of another pair"""
    from open_pcc_metric_b200 import _native as N
    vox, _ = ctxs
    A, B = _case("dense_dups_outliers")
    C = _case("lattice")[0][:3000] - 150
    a, b = vox.cloud(A), vox.cloud(B)
    vox.build_pair(a, b)
    a.close()                                   # b keeps the shared brick arrays alive
    ki, kd = b.knn_self(4)
    oi, od = cnn.knn(B, B, 4)
    assert np.array_equal(kd, od) and np.array_equal(ki, oi)
    mn, mx, per = b.self_nn_minmax(per_point=True)
    assert np.array_equal(per, np.sqrt(od[:, 1]))
    c = vox.cloud(C)
    vox.build_pair(b, c)                        # b is indexed already: c gets its own (pencil) index
    for q, s_, Q, S in ((b, c, B, C), (c, b, C, B)):
        idx, d2 = vox.nn(q, s_)
        oi, od = cnn.knn(S, Q, 1)
        assert np.array_equal(d2, od[:, 0]) and np.array_equal(idx, oi[:, 0])
    r = vox.pair_eval(b, c, 0)
    assert r.dir[0].n == len(B) and r.dir[1].n == len(C)
    b.close(); c.close()


@pytest.mark.parametrize("name", ["surface", "dense_dups_outliers", "sparse", "far_apart", "one_point_each"])
def test_split_pair_slabs_add_up_to_the_whole(ctxs, name):
    """pccm_ctx_set_shard: every rank indexes and queries its slab of z only (halo of two bricks); the partial
    results of the ranks add up to the single-GPU evaluation -- D1 sums / maxima / boundary distances bit for bit,
    float sums to rounding; queries that must look beyond the halo make that rank index the whole pair."""
    from open_pcc_metric_b200 import _native as N
    vox, _ = ctxs
    A, B = _case(name)
    n = min(len(A), len(B))
    A, B = A[:n], B[:n]
    rng = np.random.default_rng(12)
    ca, na = _attrs(rng, n)
    cb, nb = _attrs(rng, n)
    flags = N.EVAL_D2 | N.EVAL_COLOR
    a, b = vox.cloud(A, ca, na), vox.cloud(B, cb, nb)
    vox.build_pair(a, b)
    full = vox.pair_eval(a, b, flags, YUV)
    fmn, fmx, _ = (a.self_nn_minmax() if n >= 2 else (0.0, 0.0, None))
    a.close(); b.close()
    for world in (2, 5):
        parts, mm = [], []
        for r in range(world):
            vox.set_shard(r, world)
            try:
                a, b = vox.cloud(A, ca, na), vox.cloud(B, cb, nb)
                vox.build_pair(a, b)
            finally:
                vox.set_shard(0, 1)
            assert a.info().sharded == world
            parts.append(vox.pair_eval(a, b, flags, YUV, rank=r, world=world))
            if n >= 2:
                mm.append(a.self_nn_minmax()[:2])
            a.close(); b.close()
        for d in range(2):
            assert sum(p.dir[d].n for p in parts) == full.dir[d].n == n, (name, world, d)
            assert sum(p.dir[d].sum_d1_u64 for p in parts) == full.dir[d].sum_d1_u64
            assert max(p.dir[d].max_d1 for p in parts) == full.dir[d].max_d1
            assert max(p.dir[d].max_d2 for p in parts) == full.dir[d].max_d2
            assert np.isclose(sum(p.dir[d].sum_d2 for p in parts), full.dir[d].sum_d2, rtol=1e-12)
            for c in range(3):
                assert np.isclose(sum(p.dir[d].color_sum[c] for p in parts), full.dir[d].color_sum[c], rtol=1e-12)
                assert max(p.dir[d].color_max[c] for p in parts) == full.dir[d].color_max[c]
        if n >= 2:
            assert min(m[0] for m in mm) == fmn and max(m[1] for m in mm) == fmx, (name, world)


def test_host_narrowing_is_transparent(ctxs):
    """float64 HOST coordinates / colours are narrowed to uint16 / uchar by host threads while being checked (AVX2 or
    scalar); values that do not fit (a fraction, a colour that is not k / 255, NaN) make the array go up unchanged.
    Either way the evaluation is bit-identical to the one without narrowing."""
    from open_pcc_metric_b200 import _native as N
    from open_pcc_metric_b200.synth import synth_pair
    plain, _ = ctxs
    os.environ["PCCM_HOST_NARROW"] = "1"           # (opt-in: it only pays on links slower than the packing threads)
    try:
        vox = N.Context(0)
    finally:
        del os.environ["PCCM_HOST_NARROW"]
    A, B = synth_pair(10, 300_000, 5, dedup=False, oversample=4)
    n = len(A)
    assert n > 2 * (1 << 17)                       # several chunks, the last one partial
    flags = N.EVAL_D2 | N.EVAL_COLOR
    variants = {"clean": (A.points, A.colors)}
    P2 = A.points.copy(); P2[n - 3, 1] += 0.5      # a fraction in the last chunk: the pair becomes a float pair
    variants["fraction"] = (P2, A.colors)
    C2 = A.colors.copy(); C2[n // 2, 2] = 0.3      # not k / 255: float64 colour arrays
    variants["odd colour"] = (A.points, C2)
    for name, (pts, col) in variants.items():
        outs = []
        for ctx in (vox, plain):
            a, b = ctx.cloud(pts, col, A.normals), ctx.cloud(B.points, B.colors, B.normals)
            ctx.build_pair(a, b)
            assert bool(a.info().colors_u8) == (name != "odd colour")
            assert a.info().index_kind == (1 if name == "fraction" else 0) or name == "fraction"
            outs.append(bytes(ctx.pair_eval(a, b, flags, YUV)))
            a.close(); b.close()
        assert outs[0] == outs[1], name
    bad = A.points.copy(); bad[7, 0] = np.nan
    c = vox.cloud(bad)
    with pytest.raises(N.PccmError, match="NaN"):
        c.build_index()
    c.close()
    vox.close()


def test_bulk_copy_staging_is_transparent(ctxs):
    """cp.async.bulk staging (statistics pass: float64 rows; epilogue: ranks / normals / colours of a tile) against the
    plain-load builds of the same kernels (PCCM_STATS_TMA=0, PCCM_EPI_TMA=0): bit-identical evaluations -- host arrays,
    device tensors whose first element sits 8 bytes above a 16-byte line (the copy starts at the line below), point
    counts that leave partial tiles and put the second cloud's ranks at every offset within a line."""
    import torch
    from open_pcc_metric_b200 import _native as N
    staged, _ = ctxs
    os.environ["PCCM_STATS_TMA"] = "0"
    os.environ["PCCM_EPI_TMA"] = "0"
    try:
        plain = N.Context(0)
    finally:
        del os.environ["PCCM_STATS_TMA"], os.environ["PCCM_EPI_TMA"]
    A0, B0 = _case("surface")
    rng = np.random.default_rng(21)
    flags = N.EVAL_D2 | N.EVAL_COLOR

    def shifted(x):                                   # device copy whose data pointer is 8 (mod 16)
        flat = torch.empty(x.size + 3, dtype=torch.float64, device="cuda:0")
        off = 1 if flat.data_ptr() % 16 == 0 else 2
        view = flat[off:off + x.size].view(x.shape)
        view.copy_(torch.from_numpy(np.ascontiguousarray(x)))
        assert view.data_ptr() % 16 == 8
        return view

    try:
        n0 = min(len(A0), len(B0))
        assert n0 > 5000
        for n in (n0, n0 - 1, n0 - 2, n0 - 3, 513, 255):
            A, B = A0[:n], B0[:n]
            ca, na = _attrs(rng, n)
            cb, nb = _attrs(rng, n)
            outs = []
            for ctx, dev in ((plain, False), (staged, False), (staged, True)):
                f = shifted if dev else (lambda x: x)
                keep = [f(A), f(ca), f(na), f(B), f(cb), f(nb)]
                a, b = ctx.cloud(keep[0], keep[1], keep[2]), ctx.cloud(keep[3], keep[4], keep[5])
                ctx.build_pair(a, b)
                outs.append(bytes(ctx.pair_eval(a, b, flags, YUV)))
                a.close(); b.close()
            assert outs[0] == outs[1] == outs[2], n
    finally:
        plain.close()


@pytest.mark.parametrize("name", NAMES)
def test_voxel_coordinates_from_rows_equal_scattered_stores(ctxs, name):
    """The voxel coordinates of large pairs are written brick by brick from the occupancy rows (vx_rowbase_kernel),
    those of small pairs by every point in the place pass; PCCM_VXYZ_ROWS forces either.  Same evaluation bit for bit --
    whole pairs, a slab of a split pair, neighbour indices and distances, boundary distances."""
    from open_pcc_metric_b200 import _native as N
    A, B = _case(name)
    n = min(len(A), len(B))
    A, B = A[:n], B[:n]
    rng = np.random.default_rng(31)
    ca, na = _attrs(rng, n)
    cb, nb = _attrs(rng, n)
    flags = N.EVAL_D2 | N.EVAL_COLOR | N.EVAL_PERPOINT
    outs = []
    for mode in ("0", "1"):
        os.environ["PCCM_VXYZ_ROWS"] = mode
        try:
            ctx = N.Context(0)
        finally:
            del os.environ["PCCM_VXYZ_ROWS"]
        try:
            got = []
            for shard in ((0, 1), (1, 3)):
                ctx.set_shard(*shard)
                try:
                    a, b = ctx.cloud(A, ca, na), ctx.cloud(B, cb, nb)
                    ctx.build_pair(a, b)
                finally:
                    ctx.set_shard(0, 1)
                got.append(bytes(ctx.pair_eval(a, b, flags, YUV, rank=shard[0], world=shard[1])))
                if shard == (0, 1):
                    for d in range(2):
                        got.append(ctx.pair_get(N.GET_IDX, d, n).tobytes())
                        got.append(ctx.pair_get(N.GET_D2, d, n).tobytes())
                if n >= 2:
                    got.append(a.self_nn_minmax()[:2])
                a.close(); b.close()
            outs.append(got)
        finally:
            ctx.close()
    assert outs[0] == outs[1], name
