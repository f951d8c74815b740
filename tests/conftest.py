import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    # build libpccm.so, the C oracle and the CPU stepping harness once (no GPU needed to build)
    import __graft_entry__ as g
    g.build()


class Golden:
    """One fixture produced by oracle/make_golden.py (unmodified reference over the stand-in)."""

    def __init__(self, name):
        self.name = name
        with open(os.path.join(GOLDEN, name + ".json")) as f:
            self.meta = json.load(f)
        self.arr = dict(np.load(os.path.join(GOLDEN, name + ".npz")))

    def inputs(self):
        if "in_pts_a" in self.arr:
            g = self.arr.get
            return dict(pts_a=g("in_pts_a"), pts_b=g("in_pts_b"), col_a=g("in_col_a"), col_b=g("in_col_b"),
                        nrm_a=g("in_nrm_a"), nrm_b=g("in_nrm_b"))
        s = self.meta["synth"]
        from open_pcc_metric_b200 import synth
        A, B = synth.synth_pair(s["bits"], s["target_n"], synth.BASE_SEED + s["seed_offset"],
                                with_colors=s["with_colors"], with_normals=s["with_normals"])
        chk = self.arr["in_checksum"]
        assert (A.points.sum(), B.points.sum(), len(A), len(B)) == tuple(chk), "synthetic input drifted"
        return dict(pts_a=A.points, pts_b=B.points, col_a=A.colors, col_b=B.colors, nrm_a=A.normals, nrm_b=B.normals)

    def option_sets(self):
        return [json.loads(t) for t in self.meta["results"]]

    def results(self, opt):
        """{key tuple: float64 ndarray} for one option set."""
        tag = json.dumps(opt, sort_keys=True)
        out = {}
        for k, v in self.meta["results"][tag].items():
            key = tuple(json.loads(k))
            vals = v if isinstance(v, list) else [v]
            out[key] = np.array([float.fromhex(x) for x in vals]) if isinstance(v, list) else np.float64(float.fromhex(v))
        return out

    def errors(self, opt):
        tag = json.dumps(opt, sort_keys=True)
        return {tuple(json.loads(k)): v for k, v in self.meta["errors"].get(tag, {}).items()}

    def order(self, opt):
        tag = json.dumps(opt, sort_keys=True)
        return [tuple(json.loads(k)) for k in self.meta["order"][tag]]


GOLDEN_NAMES = ["ka1", "ties", "vox_small", "vox_nonormals", "float_small", "lidar_small", "identical",
                "short_b", "tiny", "config1"]


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = Golden(name)
        return cache[name]
    return get


# D1-family keys: exact integers / exact float64 given identical neighbour distances
def is_d1_family(key):
    name = key[0] if key[0] != "SymmetricMetric" else key[1]
    if name in ("MinSqrtDistance", "MaxSqrtDistance"):
        return True
    if name in ("GeoMSE", "GeoPSNR", "GeoHausdorffDistance", "GeoHausdorffDistancePSNR"):
        p2p = key[2] if key[0] != "SymmetricMetric" else key[3]
        return p2p is False
    return False


def assert_metric_close(key, got, want, rtol=1e-6, atol=1e-30, exact_d1=True):
    """exact_d1: integer (voxelised) clouds -> the D1 family is bit exact; float clouds sum in a
    different order than numpy's pairwise np.sum -> pass exact_d1=False."""
    g = np.asarray(got, dtype=np.float64)
    w = np.asarray(want, dtype=np.float64)
    assert g.shape == w.shape, (key, g.shape, w.shape)
    name = key[0] if key[0] != "SymmetricMetric" else key[1]
    if exact_d1 == "max_only" and name in ("GeoMSE", "GeoPSNR"):
        exact_d1 = False      # float clouds: the sum order differs from numpy's, the maxima do not
    if exact_d1 and is_d1_family(key):
        assert np.array_equal(g, w), (key, g, w)
    else:
        assert np.allclose(g, w, rtol=rtol, atol=atol, equal_nan=True), (key, g, w)
