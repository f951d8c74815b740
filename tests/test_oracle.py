"""CPU tier: the oracle against (a) the outputs of the UNMODIFIED reference modules
(tests/golden, made by oracle/make_golden.py), (b) the facts the reference's own tests
assert, (c) hand-derived known answers, (d) itself (numpy vs plain C)."""
import numpy as np
import pytest

from conftest import GOLDEN_NAMES, assert_metric_close
from oracle import cnn, o3d_standin as o3s, reference_port as rp


def _oracle_for(g):
    i = g.inputs()
    return rp.PairOracle(i["pts_a"], i["pts_b"], i["col_a"], i["col_b"], i["nrm_a"], i["nrm_b"])


@pytest.mark.parametrize("name", [n for n in GOLDEN_NAMES if n != "config1"])
def test_port_matches_reference_outputs(golden, name):
    g = golden(name)
    o = _oracle_for(g)
    # per-point products of cloud_pair.py:10-42 -- bit exact
    assert np.array_equal(o.d2[0], g.arr["d2_l"]) and np.array_equal(o.d2[1], g.arr["d2_r"])
    assert np.array_equal(o.idx[0], g.arr["idx_l"]) and np.array_equal(o.idx[1], g.arr["idx_r"])
    if "est_nrm_a" in g.arr:
        assert np.array_equal(o.nrm[0], g.arr["est_nrm_a"]) and np.array_equal(o.nrm[1], g.arr["est_nrm_b"])
    for opt in g.option_sets():
        want = g.results(opt)
        errs = g.errors(opt)
        peak = None
        if ("GeoPSNR", True, False) in errs:        # reference could not build the hull (e.g. 3 points)
            peak = 1.0
        n_checked = 0
        for is_left in (True, False):
            for p2p in (False, True):
                k = ("GeoMSE", is_left, p2p)
                if k in want:
                    assert_metric_close(k, o.geo_mse(is_left, p2p), want[k], rtol=1e-12)
                    n_checked += 1
                elif k in errs:
                    with pytest.raises(IndexError):
                        o.geo_mse(is_left, p2p)
                k = ("GeoHausdorffDistance", is_left, p2p)
                if k in want:
                    assert_metric_close(k, o.geo_hausdorff(is_left, p2p), want[k], rtol=1e-12)
                k = ("GeoHausdorffDistancePSNR", is_left, p2p)
                if k in want:
                    assert_metric_close(k, o.geo_hausdorff_psnr(is_left, p2p), want[k], rtol=1e-12)
                k = ("GeoPSNR", is_left, p2p)
                if k in want:
                    assert_metric_close(k, o.geo_psnr(is_left, p2p, peak), want[k], rtol=1e-12)
            if opt["color"]:
                for nm, fn in (("ColorMSE", o.color_mse), ("ColorPSNR", o.color_psnr),
                               ("ColorHausdorffDistance", o.color_hausdorff),
                               ("ColorHausdorffDistancePSNR", o.color_hausdorff_psnr)):
                    k = (nm, is_left, opt["color"])
                    assert_metric_close(k, fn(is_left, opt["color"]), want[k], rtol=1e-12, atol=1e-30)
        assert n_checked >= 2
        assert_metric_close(("MinSqrtDistance",), o.boundary_sqrt_distances()[0], want[("MinSqrtDistance",)])
        assert_metric_close(("MaxSqrtDistance",), o.boundary_sqrt_distances()[1], want[("MaxSqrtDistance",)])


def test_port_evaluate_full_dict(golden):
    """evaluate() reproduces as_dict() of the reference, key for key, including the pooled entries."""
    g = golden("vox_small")
    o = _oracle_for(g)
    for opt in g.option_sets():
        want = {k: v for k, v in g.results(opt).items() if not k[0].startswith("ColorHausdorff")}
        got = rp.evaluate(o, **opt)
        assert set(got) == set(want)
        for k in want:
            assert_metric_close(k, got[k], want[k], rtol=1e-12)


def test_config1_reference_run(golden):
    """BASELINE.json configs[0] ran through the reference's own per-point path; the batched
    oracle reproduces every squared distance, index and the D1 PSNR bit for bit."""
    g = golden("config1")
    i = g.inputs()
    idx_l, d2_l = rp.neighbour_pass(i["pts_a"], i["pts_b"])
    idx_r, d2_r = rp.neighbour_pass(i["pts_b"], i["pts_a"])
    assert np.array_equal(d2_l, g.arr["d2_l"].astype(np.float64)) and np.array_equal(idx_l, g.arr["idx_l"])
    assert np.array_equal(d2_r, g.arr["d2_r"].astype(np.float64)) and np.array_equal(idx_r, g.arr["idx_r"])
    opt = g.option_sets()[0]
    want = g.results(opt)
    peak = np.max(o3s.minimal_obb_extent(i["pts_a"]))
    for is_left, d2 in ((True, d2_l), (False, d2_r)):
        mse = np.sum(d2) / len(d2)
        assert mse == want[("GeoMSE", is_left, False)]
        assert 10 * np.log10(peak ** 2 / mse) == want[("GeoPSNR", is_left, False)]
        assert np.max(d2) == want[("GeoHausdorffDistance", is_left, False)]


def test_ka1_hand_derived(golden):
    """KA-1 = the reference's own fixture (tests/unit/test_metric.py:13-26), values derived by hand
    in SURVEY.md section 4."""
    g = golden("ka1")
    o = _oracle_for(g)
    assert np.array_equal(o.idx[0], [0, 1, 2]) and np.array_equal(o.idx[1], [0, 1, 2])
    assert np.allclose(o.d2[0], 0.14) and np.allclose(o.d2[1], 0.14)
    assert np.max(o.d2[0]) == 0.14000000000000004
    assert np.allclose(o.geo_mse(True, False), 0.14)
    assert np.allclose(o.boundary_sqrt_distances(), (np.sqrt(2), np.sqrt(2)))
    assert np.allclose(o.color_mse(True, "rgb"), [0.01, 0.04, 0.09])
    assert np.allclose(o.color_mse(True, "ycc"), [0.03458112, 0.00377733, 0.00297898], rtol=1e-6)
    assert np.allclose(o.color_mse(True, "yuv")[:2], [0.04, 0.04])
    assert abs(o.color_mse(True, "yuv")[2]) < 1e-30
    assert np.allclose(o.color_psnr(True, "rgb"), [68.1308, 62.1102, 58.5884], atol=1e-4)


def test_reference_own_assertions():
    """What tests/unit/test_metric.py:30-70 pins: row norm of ones = sqrt(3); D1 passes the squared
    distances through; D2 squares the plane error."""
    e = np.ones((5, 3))
    assert np.allclose(np.sqrt((e * e).sum(1)), np.sqrt(3))
    o = rp.PairOracle(np.eye(3), np.eye(3) + 0.1)
    assert o.euclidean_distance(True, False) is o.d2[0]
    assert np.array_equal(o.euclidean_distance(True, True), np.square(o.plane_errors(True)))


def test_c_oracle_matches_numpy():
    rng = np.random.default_rng(0)
    P = rng.integers(0, 40, (3000, 3)).astype(float)   # ties everywhere
    Q = rng.integers(0, 40, (1500, 3)).astype(float)
    for k in (1, 2, 30):
        i1, d1 = cnn.knn(P, Q, k)
        i2, d2 = o3s._brute_knn(P, Q, k)
        assert np.array_equal(i1, i2) and np.array_equal(d1, d2)
    Pf = rng.random((20000, 3))                          # tree-assisted path of exact_knn
    i3, d3 = o3s.exact_knn(Pf, Pf[:4000], 30)
    i4, d4 = cnn.knn(Pf, Pf[:4000], 30)
    assert np.array_equal(i3, i4) and np.array_equal(d3, d4)
    nn, _ = cnn.knn(Pf[:3000], Pf[:3000], 30)
    assert np.array_equal(cnn.normals(Pf[:3000], nn), o3s.estimate_normals_array(Pf[:3000], 30, nn))


def test_normal_estimation_degenerate_branches():
    # < 3 neighbours -> identity covariance -> (0, 0, 1)
    assert np.array_equal(o3s.estimate_normals_array(np.array([[0., 0, 0], [1, 0, 0]])), [[0, 0, 1], [0, 0, 1]])
    # identical points -> zero covariance -> zero vector -> (0, 0, 1)
    assert np.array_equal(o3s.estimate_normals_array(np.zeros((4, 3))), np.tile([0., 0, 1], (4, 1)))
    # axis-aligned planes: diagonal covariance picks the axis with the smallest variance
    g = np.stack(np.meshgrid(np.arange(5.), np.arange(5.), indexing="ij"), -1).reshape(-1, 2)
    for axis in range(3):
        pts = np.insert(g, axis, 7.0, axis=1)
        n = o3s.estimate_normals_array(pts, 30)
        assert np.allclose(np.abs(n[:, axis]), 1.0) and np.allclose(np.delete(n, axis, 1), 0.0)
    # generic plane: normal is orthogonal to it
    rng = np.random.default_rng(1)
    uv = rng.random((200, 2))
    pts = uv[:, :1] * np.array([1., 2, 0.5]) + uv[:, 1:] * np.array([-1., 0.3, 2])
    n = o3s.estimate_normals_array(pts, 30)
    w = np.cross([1., 2, 0.5], [-1., 0.3, 2])
    w /= np.linalg.norm(w)
    assert np.allclose(np.abs(n @ w), 1.0, atol=1e-9)


def test_minimal_obb_known_box():
    rng = np.random.default_rng(2)
    box = rng.random((500, 3)) * np.array([4.0, 2.0, 1.0])
    corners = np.array([[x, y, z] for x in (0, 4.) for y in (0, 2.) for z in (0, 1.)])
    th = 0.7
    R = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]])
    pts = np.concatenate([box, corners]) @ R.T + 5.0
    ext = o3s.minimal_obb_extent(pts)
    assert np.allclose(sorted(ext), [1.0, 2.0, 4.0], atol=1e-9)


def test_tie_average_fixture_matches_the_restatement(golden):
    """tests/golden/tie_average.json (oracle/make_golden_ties.py): the brute-force restatement of the tie-averaged
    mode reproduces its committed values bit for bit, and differs from the single-neighbour values exactly where ties exist."""
    import json
    import os
    from conftest import GOLDEN
    from oracle import reference_port as rp
    fx = json.load(open(os.path.join(GOLDEN, "tie_average.json")))
    for name, rec in fx.items():
        i = golden(name).inputs()
        n = rec["n"]
        rng = np.random.default_rng(rec["seed"])
        nrm = [rng.normal(0, 1, (n, 3)) for _ in range(2)]
        nrm = [v / np.linalg.norm(v, axis=1, keepdims=True) for v in nrm]
        col = [rng.integers(0, 256, (n, 3)).astype(np.float64) / 255.0 for _ in range(2)]
        o = rp.PairOracle(i["pts_a"][:n], i["pts_b"][:n], col[0], col[1], nrm[0], nrm[1])
        for is_left, key in ((True, "left"), (False, "right")):
            pe2, cd2 = rp.tie_average(o, is_left, "yuv")
            assert float(pe2.sum()).hex() == rec[key]["sum_d2"] and float(pe2.max()).hex() == rec[key]["max_d2"]
            assert [float(x).hex() for x in cd2.sum(0)] == rec[key]["color_sum"]
        if name == "ties":       # the lattice fixture has ties: averaging must change the plane error
            q, s = o.pts[0], o.pts[1]
            single = (o.error_vector(True) * o.nrm[1][o.idx[0]]).sum(1) ** 2
            assert not np.allclose(single.sum(), rp.tie_average(o, True)[0].sum())
