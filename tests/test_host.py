"""CPU tier: the C-ABI library loads and exports everything include/pccm.h declares, the
host-side mirror of the reference interface behaves like the reference, and nothing
computes without a GPU."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT, assert_metric_close
from open_pcc_metric_b200 import _native as N
from open_pcc_metric_b200 import metric as M
from open_pcc_metric_b200 import obb
from open_pcc_metric_b200.calculator import CalculateResult, MetricCalculator
from open_pcc_metric_b200.io import read_point_cloud, write_ply
from open_pcc_metric_b200.options import CalculateOptions, transform_options


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "pccm.h")).read()
    declared = set(re.findall(r"\b(pccm_[a-z0-9_]+)\s*\(", header))
    assert declared == set(N.EXPORTS), declared ^ set(N.EXPORTS)
    L = N.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.pccm_version() == 100


def test_struct_layouts_match_header():
    import ctypes
    assert ctypes.sizeof(N.DirResult) == 8 * 3 + 4 * 2 + 8 * 4 + 8 * 6
    assert ctypes.sizeof(N.PairResult) == 2 * ctypes.sizeof(N.DirResult)
    assert ctypes.sizeof(N.CloudInfo) == 8 + 4 * 10 + 8 + 8 * 6
    assert ctypes.sizeof(N.Timings) == 13 * 8 + 7 * 8   # 13 stage timers, 7 counters (include/pccm.h)


def test_no_cpu_fallback():
    """Without a GPU the product refuses to compute (it never routes through the oracle)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(N.PccmError, match="no CUDA device"):
        N.Context(0)
    from open_pcc_metric_b200.cloud_pair import CloudPair
    from open_pcc_metric_b200.synth import Cloud
    with pytest.raises(N.PccmError):
        CloudPair(Cloud(np.zeros((4, 3))), Cloud(np.ones((4, 3))))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "open_pcc_metric_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "oracle/" not in src or f.endswith(".md"), f


# ---- the reference's own unit tests, re-stated against this package --------------------
@pytest.mark.parametrize("is_left", [True, False])
def test_default_error_vector(is_left):
    ev = M.ErrorVector(is_left=is_left, point_to_plane=False)
    pev = M.PrimaryErrorVector(is_left=is_left)
    pev.value = np.ones((5, 3), dtype="float64")
    ev.calculate(pev)
    assert np.allclose(ev.value, np.sqrt(3) * np.ones(5))


@pytest.mark.parametrize("is_left,point_to_plane", [(True, False), (False, False), (True, True), (False, True)])
def test_default_euclidean_distance(is_left, point_to_plane):
    ed = M.EuclideanDistance(is_left=is_left, point_to_plane=point_to_plane)
    pev = M.PrimaryErrorVector(is_left=is_left)
    pev.value = 2 * np.ones((5,))
    nd = M.NeighbourDistances(is_left=is_left)
    nd.value = 4 * np.ones((5,))
    ed.calculate(nd, pev)
    assert np.allclose(nd.value, ed.value)


# ---- metric graph over a stub pair (pure host arithmetic) ---------------------------------
class StubPair:
    """Duck-typed CloudPair built from oracle-free hand data: exercises the dependency
    graph path (no fused results available)."""

    def __init__(self):
        rng = np.random.default_rng(0)
        self.A = rng.integers(0, 9, (12, 3)).astype(float)
        self.B = rng.integers(0, 9, (12, 3)).astype(float)
        self.ca, self.cb = rng.random((12, 3)), rng.random((12, 3))
        self.na = np.tile([0, 0, 1.0], (12, 1))
        self.nb = np.tile([0, 1.0, 0], (12, 1))
        d = ((self.A[:, None] - self.B[None]) ** 2).sum(-1)
        self.il, self.ir = d.argmin(1), d.argmin(0)

        class C:
            pass
        self.clouds = (C(), C())
        self.clouds[0].normals, self.clouds[1].normals = self.na, self.nb

    def get_left_error_vector(self): return self.A - self.B[self.il]
    def get_right_error_vector(self): return self.B - self.A[self.ir]
    def get_left_neighbour_distances(self): return ((self.A - self.B[self.il]) ** 2).sum(1)
    def get_right_neighbour_distances(self): return ((self.B - self.A[self.ir]) ** 2).sum(1)
    def get_boundary_sqrt_distances(self): return np.array([1.0, 2.0, 3.0])
    def get_extent(self): return np.array([3.0, 9.0, 2.0])
    def get_left_colors(self): return self.ca
    def get_right_colors(self): return self.cb
    def get_left_neighbour_colors(self): return self.cb[self.il]
    def get_right_neighbour_colors(self): return self.ca[self.ir]


def test_metric_graph_values():
    p = StubPair()
    res = MetricCalculator(p).calculate(transform_options(CalculateOptions("ycc", True, True))).as_dict()
    dl = p.get_left_neighbour_distances()
    assert res[("GeoMSE", True, False)] == dl.sum() / 12
    assert res[("GeoPSNR", True, False)] == 10 * np.log10(81.0 / (dl.sum() / 12))
    assert res[("GeoHausdorffDistance", True, False)] == dl.max()
    assert res[("MinSqrtDistance",)] == 1.0 and res[("MaxSqrtDistance",)] == 3.0
    assert res[("GeoHausdorffDistancePSNR", True, False)] == 10 * np.log10(9.0 / dl.max())
    # D2, quirk Q1: left uses the RIGHT cloud's normals at the query index
    pe = (p.get_left_error_vector() * p.nb).sum(1)
    assert np.allclose(res[("GeoMSE", True, True)], (pe ** 2).sum() / 12, rtol=1e-14)
    T = np.array([[0.2126, 0.7152, 0.0722], [-0.1146, -0.3854, 0.5], [0.5, -0.4542, -0.0458]])
    diff = p.ca @ T.T - p.cb[p.il] @ T.T
    assert np.allclose(res[("ColorMSE", True, "ycc")], (diff ** 2).mean(0), rtol=1e-13)
    assert np.allclose(res[("ColorPSNR", True, "ycc")], 10 * np.log10(1.0 / (diff ** 2).mean(0)), rtol=1e-13)
    # pooled: error -> larger norm, PSNR -> smaller norm (metric.py:475-485)
    l, r = res[("GeoMSE", True, False)], res[("GeoMSE", False, False)]
    assert res[("SymmetricMetric", "GeoMSE", True, False, "GeoMSE", False, False)] == max(l, r)
    l, r = res[("GeoPSNR", True, False)], res[("GeoPSNR", False, False)]
    assert res[("SymmetricMetric", "GeoPSNR", True, False, "GeoPSNR", False, False)] == min(l, r)


def test_color_hausdorff_rgb_scale_quirk():
    p = StubPair()
    calc = MetricCalculator(p)
    hd = calc._metric_recursive_calculate(M.ColorHausdorffDistance(True, "rgb")).value
    assert np.allclose(hd, ((255 * (p.ca - p.cb[p.il])) ** 2).max(0))
    hd = calc._metric_recursive_calculate(M.ColorHausdorffDistance(True, "yuv")).value
    T = np.array([[0.25, 0.5, 0.25], [1, 0, -1], [-0.5, 1, -0.5]])
    assert np.allclose(hd, ((p.ca @ T.T - p.cb[p.il] @ T.T) ** 2).max(0))
    psnr = calc._metric_recursive_calculate(M.ColorHausdorffDistancePSNR(True, "rgb")).value
    assert np.allclose(psnr, 10 * np.log10(255.0 ** 2 / ((255 * (p.ca - p.cb[p.il])) ** 2).max(0)))


def test_d2_index_error_when_other_cloud_is_shorter():
    ev = M.ErrorVector(True, True)
    pev = M.PrimaryErrorVector(True)
    pev.value = np.ones((5, 3))
    cn = M.CloudNormals(False)
    cn.value = np.ones((3, 3))
    with pytest.raises(IndexError):
        ev.calculate(pev, cn)


def test_options_order_matches_reference(golden):
    """The metric list has the reference's content and ORDER (options.py:32-174); the order was
    recorded from the unmodified reference by oracle/make_golden.py."""
    g = golden("vox_small")
    for opt in g.option_sets():
        want = [k for k in g.order(opt) if not k[0].startswith("ColorHausdorff")]
        got = [m._key() for m in transform_options(CalculateOptions(**opt))]
        assert got == want
    assert len(transform_options(CalculateOptions())) == 8
    assert len(transform_options(CalculateOptions(color="rgb"))) == 14
    assert len(transform_options(CalculateOptions(hausdorff=True))) == 14
    assert len(transform_options(CalculateOptions("ycc", True, True))) == 32


def test_memo_is_per_instance():
    """Quirk Q2 fixed on purpose: a second calculator must not see the first pair's values."""
    p1, p2 = StubPair(), StubPair()
    p2.B = p2.B + 1
    v1 = MetricCalculator(p1).calculate([M.GeoMSE(True, False)]).as_dict()[("GeoMSE", True, False)]
    v2 = MetricCalculator(p2).calculate([M.GeoMSE(True, False)]).as_dict()[("GeoMSE", True, False)]
    assert v1 != v2
    c = MetricCalculator(p1)
    a = c._metric_recursive_calculate(M.GeoMSE(True, False))
    assert c._metric_recursive_calculate(M.GeoMSE(True, False)) is a   # memo within one instance


def test_symmetric_metric_validation():
    with pytest.raises(ValueError):
        M.SymmetricMetric([M.GeoMSE(True, False)], False)
    with pytest.raises(ValueError):
        M.SymmetricMetric([M.GeoMSE(True, False), M.GeoPSNR(False, False)], False)

    class Weird(M.AbstractMetric):
        def calculate(self):
            pass
    with pytest.raises(RuntimeError):
        MetricCalculator(StubPair())._metric_recursive_calculate(Weird())


def test_result_table_format():
    p = StubPair()
    res = MetricCalculator(p).calculate(transform_options(CalculateOptions()))
    df = res.as_df()
    assert list(df.columns) == ["label", "is_left", "point-to-plane", "value"]
    assert df["label"].tolist() == ["MinSqrtDistance", "MaxSqrtDistance", "GeoMSE", "GeoMSE", "GeoMSE(symmetric)",
                                    "GeoPSNR", "GeoPSNR", "GeoPSNR(symmetric)"]
    assert df["is_left"].tolist()[:5] == ["", "", True, False, ""]
    assert isinstance(res, CalculateResult) and "GeoMSE" in str(res)


def test_peak_helpers():
    assert obb.aabb_diag([0, 0, 0], [3, 4, 12]) == 13.0
    assert obb.resolution_peak([1023, 5, 7]) == 1023.0 and obb.resolution_peak([1, 1, 1], 12) == 4095.0


@pytest.mark.parametrize("binary", [True, False])
def test_ply_roundtrip(tmp_path, binary):
    from open_pcc_metric_b200.geometry import PointCloud
    rng = np.random.default_rng(6)
    pc = PointCloud(rng.integers(0, 1024, (50, 3)).astype(float),
                    rng.integers(0, 256, (50, 3)) / 255.0, rng.normal(0, 1, (50, 3)))
    path = str(tmp_path / "c.ply")
    write_ply(path, pc, binary=binary)
    back = read_point_cloud(path)
    assert np.array_equal(back.points, pc.points)
    assert np.array_equal(back.colors, pc.colors)       # uchar / 255.0, exactly Open3D's convention
    assert np.allclose(back.normals, pc.normals, rtol=1e-15)
    assert back.has_colors() and back.has_normals()
    xyz = str(tmp_path / "c.xyz")
    np.savetxt(xyz, np.asarray(pc.points))
    assert np.array_equal(read_point_cloud(xyz).points, pc.points)


def _load_bench():
    import importlib.util
    spec = importlib.util.spec_from_file_location("pccm_bench", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_bench_synthetic_cache_is_atomic_and_self_healing(monkeypatch):
    """The ranks of one torchrun job generate the split pair at the same moment: only rank 0 keeps it (temporary file +
    rename), the other ranks never write, and a damaged cache file is regenerated instead of crashing the run."""
    import argparse
    import glob
    bench = _load_bench()
    args = argparse.Namespace(config="split", points=12_000, mode="auto")
    pat = "/tmp/pccm_bench_vox12_12000_*"
    for f in glob.glob(pat):
        os.remove(f)
    try:
        w = bench.Workload(args, 2)
        monkeypatch.setenv("RANK", "1")
        a1, b1 = w.gen(1)
        assert glob.glob(pat) == []                                   # rank 1 keeps nothing
        monkeypatch.setenv("RANK", "0")
        a0, b0 = w.gen(0)
        files = glob.glob(pat)
        assert len(files) == 1 and not files[0].endswith(".tmp.npz")
        a2, b2 = w.gen(0)                                             # read back from the cache
        for x, y in ((a0, a1), (a0, a2), (b0, b1), (b0, b2)):
            assert np.array_equal(x.points, y.points) and np.array_equal(x.colors, y.colors) and np.array_equal(x.normals, y.normals)
        with open(files[0], "wb") as fh:
            fh.write(b"not a zip file")
        a3, _ = w.gen(0)
        assert np.array_equal(a3.points, a0.points)
        assert np.array_equal(np.load(files[0])["p"], a0.points)      # ... and repaired
    finally:
        for f in glob.glob(pat):
            os.remove(f)


def test_reference_arm_bounded_sample_scales_to_the_pair():
    """bench.py --impl reference on the 10 M-point pair times a bounded sample per step (trees built once, the first
    queries of each direction on all cores) and scales it to the pair; on a small pair the estimate agrees with the
    full evaluation to within the noise of a sub-second measurement, and the bookkeeping is exact."""
    from oracle import cpu_baseline as cb
    from open_pcc_metric_b200 import synth
    A, B = synth.synth_pair(9, 60_000, 3, dedup=False, oversample=4)
    pre = cb.build_trees(A, B)
    full = cb.cpu_best(A, B, "yuv", True)
    samp = cb.cpu_best(A, B, "yuv", True, 20_000, pre)
    assert samp["sampled"] is True and samp["sample_queries"] == 40_000
    assert samp["queries"] == full["queries"] == len(A) + len(B)
    assert samp["seconds"] > samp["build_seconds"] == pre[1]
    assert 0.2 < samp["seconds"] / full["seconds"] < 5.0


def _write_typed_ply(path, fields, rec, fmt):
    names = {"f4": "float", "f8": "double", "u1": "uchar", "u2": "ushort", "i4": "int"}
    header = ["ply", f"format {fmt} 1.0", "comment typed test file", f"element vertex {len(rec)}"]
    header += [f"property {names[t]} {k}" for k, t in fields]
    header.append("end_header")
    with open(path, "wb") as f:
        f.write(("\n".join(header) + "\n").encode("ascii"))
        if fmt == "ascii":
            for r in rec:
                f.write((" ".join(repr(v.item()) for v in r) + "\n").encode("ascii"))
        else:
            end = "<" if fmt == "binary_little_endian" else ">"
            f.write(rec.astype(np.dtype([(k, end + t) for k, t in fields])).tobytes())


@pytest.mark.parametrize("fmt", ["binary_little_endian", "binary_big_endian", "ascii"])
@pytest.mark.parametrize("xyz_t", ["f4", "u2", "i4", "f8"])
def test_native_reader_keeps_file_types_and_the_float64_view(tmp_path, fmt, xyz_t):
    """read_point_cloud(native=True): the arrays CloudPair uploads keep the file's scalar types (a third of the bytes of
    the float64 copy on a voxelised PLY), the getters give exactly the float64 arrays of the default reader."""
    from open_pcc_metric_b200.cloud_pair import _upload
    from open_pcc_metric_b200.io import FileCloud
    rng = np.random.default_rng(17)
    n = 40
    fields = [(k, xyz_t) for k in "xyz"] + [(k, "u1") for k in ("red", "green", "blue")] + [(k, "f4") for k in ("nx", "ny", "nz")]
    rec = np.zeros(n, dtype=np.dtype([(k, t) for k, t in fields]))
    for k in "xyz":
        rec[k] = rng.integers(0, 1024, n) if xyz_t != "f4" and xyz_t != "f8" else rng.random(n).astype(xyz_t) * 100
    for k in ("red", "green", "blue"):
        rec[k] = rng.integers(0, 256, n)
    for k in ("nx", "ny", "nz"):
        rec[k] = rng.normal(0, 1, n).astype("f4")
    path = str(tmp_path / "typed.ply")
    _write_typed_ply(path, fields, rec, fmt)
    ref = read_point_cloud(path)
    got = read_point_cloud(path, native=True, pinned=(xyz_t == "u2"))    # (page-locked where CUDA is there, ordinary arrays here)
    assert isinstance(got, FileCloud) and len(got) == n and got.has_colors() and got.has_normals()
    text = fmt == "ascii"
    assert got.raw_points.dtype == np.dtype("f8" if text and xyz_t == "f4" else xyz_t)   # text keeps integer types only
    assert got.raw_colors.dtype == np.uint8
    assert got.raw_normals.dtype == np.dtype("f8" if text else "f4")
    assert got.raw_points.shape == got.raw_colors.shape == got.raw_normals.shape == (n, 3)
    assert got.raw_points.flags.c_contiguous and got.raw_points.dtype.isnative
    for name in ("points", "colors", "normals"):
        assert np.array_equal(getattr(got, name), np.asarray(getattr(ref, name))), name
        assert getattr(got, name).dtype == np.float64
        assert _upload(got, name) is getattr(got, "raw_" + name)
        assert _upload(ref, name) is getattr(ref, name)
    est = rng.normal(0, 1, (n, 3))
    got.normals = est                                    # what CloudPair does with estimated normals
    assert np.array_equal(got.normals, est) and _upload(got, "normals").dtype == np.float64


def test_native_reader_falls_back_to_float64_on_mixed_types(tmp_path):
    fields = [("x", "f4"), ("y", "f8"), ("z", "f4"), ("red", "f4"), ("green", "f4"), ("blue", "f4")]
    rec = np.zeros(5, dtype=np.dtype(fields))
    for k, _ in fields:
        rec[k] = np.arange(5) / 8.0
    path = str(tmp_path / "mixed.ply")
    _write_typed_ply(path, fields, rec, "binary_little_endian")
    got = read_point_cloud(path, native=True)
    assert got.raw_points.dtype == np.float64 and got.raw_colors.dtype == np.float64 and got.raw_normals is None
    assert got.normals is None and not got.has_normals()
    assert np.array_equal(got.colors, np.asarray(read_point_cloud(path).colors))     # float colours are kept as they are
