"""Generate tests/golden/* by running the UNMODIFIED reference package
(/root/reference/open_pcc_metric: cloud_pair.py, metric.py, calculator.py,
options.py) over the Open3D stand-in (oracle/o3d_standin.py).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Run in the build container only:

    python -m oracle.make_golden            # all fixtures
    python -m oracle.make_golden ka1 ties   # some

/root/reference does not exist on the GPU box, so the outputs are committed.
Floats are stored as C99 hex strings (bit exact); arrays go to a compressed .npz.
The reference's class-level memo (calculator.py:60, quirk Q2) is cleared before
each evaluation -- otherwise every fixture after the first would return the
first fixture's values.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE_ROOT = "/root/reference"


def _import_reference():
    from oracle import o3d_standin as o3s
    o3d = o3s.install_fake_open3d()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import logging
    import open_pcc_metric.metric as opmm  # noqa: F401  (unmodified reference)
    from open_pcc_metric.cloud_pair import CloudPair
    from open_pcc_metric.calculator import MetricCalculator
    from open_pcc_metric.options import CalculateOptions, transform_options
    logging.getLogger().handlers.clear()
    logging.getLogger().setLevel(logging.WARNING)
    return o3d, opmm, CloudPair, MetricCalculator, CalculateOptions, transform_options


def hexify(v):
    if isinstance(v, tuple):
        return [hexify(x) for x in v]
    a = np.asarray(v, dtype=np.float64)
    if a.ndim == 0:
        return float(a).hex()
    return [float(x).hex() for x in a.ravel()]


def key_str(k):
    return json.dumps(list(k))


# --------------------------------------------------------------------------
# fixtures (inputs only)
# --------------------------------------------------------------------------
def fx_ka1():
    """The reference's own deterministic fixture, tests/unit/test_metric.py:13-26."""
    n = 3
    A = np.eye(n, dtype="float64")
    err = 1e-1 * np.linspace(1.0, n, n)
    B = A + err
    return dict(pts_a=A, pts_b=B, col_a=np.copy(A), col_b=np.copy(A) + err, extent=False)


def fx_ties():
    """Integer lattice slabs with massive distance ties and duplicated points."""
    rng = np.random.default_rng(7)
    g = np.stack(np.meshgrid(np.arange(7), np.arange(6), np.arange(3), indexing="ij"), -1).reshape(-1, 3)
    A = g[rng.permutation(len(g))].astype(np.float64)
    B = (g[::2] * 1.0 + np.array([1.0, 0.0, 1.0]))
    B = np.concatenate([B, B[:9]])  # exact duplicates in the search cloud
    B = B[rng.permutation(len(B))]
    ca = rng.integers(0, 256, A.shape).astype(np.float64) / 255.0
    cb = rng.integers(0, 256, B.shape).astype(np.float64) / 255.0
    return dict(pts_a=A, pts_b=B, col_a=ca, col_b=cb)


def fx_vox_small():
    """Voxelised surface, colours, normals GIVEN (config-2 shape at test scale)."""
    from open_pcc_metric_b200 import synth
    A, B = synth.synth_pair(7, 3000, synth.BASE_SEED + 2)
    # D2 in reference mode needs N_other >= N_query in both directions (quirk Q1):
    # swap roles so that the longer cloud is searched ... impossible for both, so
    # make the clouds equally long by truncating A's tail (still shuffled order).
    n = min(len(A), len(B))
    return dict(pts_a=A.points[:n], pts_b=B.points[:n], col_a=A.colors[:n], col_b=B.colors[:n],
                nrm_a=A.normals[:n], nrm_b=B.normals[:n])


def fx_vox_nonormals():
    """Voxelised surface without normals -> k=30 estimation (config-3 shape, small)."""
    from open_pcc_metric_b200 import synth
    A, B = synth.synth_pair(7, 2500, synth.BASE_SEED + 3, with_colors=False, with_normals=False)
    n = min(len(A), len(B))
    return dict(pts_a=A.points[:n], pts_b=B.points[:n])


def fx_float_small():
    """Generic float64 coordinates and colours, no normals."""
    rng = np.random.default_rng(11)
    t = rng.random((1200, 2))
    A = np.stack([t[:, 0] * 4, t[:, 1] * 3, np.sin(t[:, 0] * 5) * np.cos(t[:, 1] * 4)], 1)
    A += rng.normal(0, 0.01, A.shape)
    B = A[rng.permutation(len(A))] + rng.normal(0, 0.02, A.shape)
    ca = rng.random(A.shape)
    cb = np.clip(ca[rng.permutation(len(A))] + rng.normal(0, 0.05, A.shape), 0, 1)
    return dict(pts_a=A, pts_b=B, col_a=ca, col_b=cb)


def fx_lidar_small():
    """float32-valued LiDAR-style scene (config-5 shape at test scale)."""
    from open_pcc_metric_b200 import synth
    A, B = synth.synth_lidar(4000, synth.BASE_SEED + 5)
    n = min(len(A), len(B))
    return dict(pts_a=A.points[:n].astype(np.float64), pts_b=B.points[:n].astype(np.float64))


def fx_identical():
    """A == B: zero error, PSNR = inf (quirk Q14)."""
    rng = np.random.default_rng(3)
    A = rng.integers(0, 32, (300, 3)).astype(np.float64)
    A = np.unique(A, axis=0)
    A = A[rng.permutation(len(A))]
    return dict(pts_a=A, pts_b=A.copy(), col_a=np.full(A.shape, 0.5), col_b=np.full(A.shape, 0.5))


def fx_short_b():
    """N_B < N_A: D1 works, D2 raises IndexError in the left direction (quirk Q1)."""
    rng = np.random.default_rng(5)
    A = rng.integers(0, 24, (400, 3)).astype(np.float64)
    B = A[:250] + rng.integers(-1, 2, (250, 3))
    return dict(pts_a=A, pts_b=B, expect_p2plane_error=True)


def fx_tiny():
    """Degenerate sizes: 5 points vs 4 (normal estimation with < 30 neighbours)."""
    A = np.array([[0, 0, 0], [4, 0, 0], [0, 4, 0], [0, 0, 4], [4, 4, 4]], dtype=np.float64)
    B = np.array([[1, 0, 0], [4, 1, 0], [0, 4, 1], [1, 1, 5], [3, 4, 4]], dtype=np.float64)
    return dict(pts_a=A, pts_b=B)


def fx_config1():
    """BASELINE.json configs[0]: vox10 ~100k vs quantised+jittered copy, D1 PSNR, run
    through the reference's own per-point path."""
    from open_pcc_metric_b200 import synth
    A, B = synth.synth_pair(10, 100_000, synth.BASE_SEED + 1, with_colors=False, with_normals=True)
    return dict(pts_a=A.points, pts_b=B.points, nrm_a=A.normals, nrm_b=B.normals,
                store_inputs=False, option_sets=[dict(color=None, hausdorff=True, point_to_plane=False)],
                synth=dict(fn="synth_pair", bits=10, target_n=100_000, seed_offset=1,
                           with_colors=False, with_normals=True))


FIXTURES = {
    "ka1": fx_ka1, "ties": fx_ties, "vox_small": fx_vox_small, "vox_nonormals": fx_vox_nonormals,
    "float_small": fx_float_small, "lidar_small": fx_lidar_small, "identical": fx_identical,
    "short_b": fx_short_b, "tiny": fx_tiny, "config1": fx_config1,
}


# --------------------------------------------------------------------------
def run_fixture(name, fx):
    o3d, opmm, CloudPair, MetricCalculator, CalculateOptions, transform_options = _import_reference()
    t0 = time.time()
    clouds = []
    for side in ("a", "b"):
        c = o3d.geometry.PointCloud()
        c.points = o3d.utility.Vector3dVector(fx["pts_" + side])
        if fx.get("col_" + side) is not None:
            c.colors = o3d.utility.Vector3dVector(fx["col_" + side])
        if fx.get("nrm_" + side) is not None:
            c.normals = o3d.utility.Vector3dVector(fx["nrm_" + side])
        clouds.append(c)
    pair = CloudPair(clouds[0], clouds[1])  # unmodified cloud_pair.py:54-80
    arrays = {}
    if fx.get("store_inputs", True):
        for k in ("pts_a", "pts_b", "col_a", "col_b", "nrm_a", "nrm_b"):
            if fx.get(k) is not None:
                arrays["in_" + k] = np.asarray(fx[k], dtype=np.float64)
    # per-point products of the hot path
    arrays["d2_l"] = np.asarray(pair.get_left_neighbour_distances())
    arrays["d2_r"] = np.asarray(pair.get_right_neighbour_distances())
    ev_l = np.asarray(pair.get_left_error_vector())
    ev_r = np.asarray(pair.get_right_error_vector())
    # recover the matched indices from the neighbour clouds via the stand-in
    from oracle.reference_port import neighbour_pass
    arrays["idx_l"], _ = neighbour_pass(np.asarray(clouds[0].points), np.asarray(clouds[1].points))
    arrays["idx_r"], _ = neighbour_pass(np.asarray(clouds[1].points), np.asarray(clouds[0].points))
    assert np.array_equal(ev_l, np.asarray(clouds[0].points) - np.asarray(clouds[1].points)[arrays["idx_l"]])
    assert np.array_equal(ev_r, np.asarray(clouds[1].points) - np.asarray(clouds[0].points)[arrays["idx_r"]])
    if fx.get("nrm_a") is None:
        arrays["est_nrm_a"] = np.asarray(clouds[0].normals)
    if fx.get("nrm_b") is None:
        arrays["est_nrm_b"] = np.asarray(clouds[1].normals)
    if fx.get("store_inputs", True) is False:
        # big fixture: keep only compact integer forms
        arrays["d2_l"] = arrays["d2_l"].astype(np.uint32)
        arrays["d2_r"] = arrays["d2_r"].astype(np.uint32)
        arrays["idx_l"] = arrays["idx_l"].astype(np.int32)
        arrays["idx_r"] = arrays["idx_r"].astype(np.int32)
        arrays["in_checksum"] = np.array([fx["pts_a"].sum(), fx["pts_b"].sum(),
                                          len(fx["pts_a"]), len(fx["pts_b"])])

    meta = {"name": name, "n_a": int(len(fx["pts_a"])), "n_b": int(len(fx["pts_b"])),
            "has_colors": fx.get("col_a") is not None, "normals_given": fx.get("nrm_a") is not None,
            "synth": fx.get("synth"), "results": {}, "errors": {}, "order": {}}
    has_col = fx.get("col_a") is not None
    option_sets = fx.get("option_sets")
    if option_sets is None:
        option_sets = [dict(color=None, hausdorff=True, point_to_plane=True)]
        if has_col:
            option_sets = [dict(color=s, hausdorff=True, point_to_plane=True) for s in ("rgb", "ycc", "yuv")]
    for opt in option_sets:
        MetricCalculator._calculated_metrics.clear()  # quirk Q2
        calc = MetricCalculator(pair)
        metrics = transform_options(CalculateOptions(**opt))
        if opt["color"] is not None:  # public but unreachable from options.py (quirk Q15)
            for is_left in (True, False):
                metrics.append(opmm.ColorHausdorffDistance(is_left, opt["color"]))
                metrics.append(opmm.ColorHausdorffDistancePSNR(is_left, opt["color"]))
        tag = json.dumps(opt, sort_keys=True)
        meta["order"][tag] = [key_str(m._key()) for m in metrics]
        res = {}
        errs = {}
        with np.errstate(all="ignore"):
            for m in metrics:
                try:
                    done = calc._metric_recursive_calculate(m)  # calculator.py:65-95
                    res[key_str(done._key())] = hexify(done.value)
                except Exception as e:  # record reference failure modes
                    errs[key_str(m._key())] = type(e).__name__
        meta["results"][tag] = res
        if errs:
            meta["errors"][tag] = errs
    meta["seconds"] = round(time.time() - t0, 2)
    os.makedirs(GOLDEN, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **arrays)
    with open(os.path.join(GOLDEN, name + ".json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print(f"[golden] {name}: n_a={meta['n_a']} n_b={meta['n_b']} "
          f"{sum(len(v) for v in meta['results'].values())} values, "
          f"{sum(len(v) for v in meta['errors'].values())} errors, {meta['seconds']} s")


def main(argv):
    names = argv or list(FIXTURES)
    for name in names:
        run_fixture(name, FIXTURES[name]())


if __name__ == "__main__":
    main(sys.argv[1:])
