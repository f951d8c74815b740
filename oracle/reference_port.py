"""Vectorised numpy restatement of the reference hot path (TEST INFRASTRUCTURE --
see oracle/__init__.py; parity unpinned at the Open3D boundary).

Every function cites the reference lines it follows (paths relative to
/root/reference/open_pcc_metric/).  Open3D calls are replaced by the stand-ins in
``o3d_standin.py``.  The port is checked bit for bit (integer / d2 quantities) or to
1e-12 (dot-product and colour sums, where numpy's own summation order differs
between the per-row loops of the reference and the vectorised forms used here)
against the outputs of the UNMODIFIED reference modules run over the same
stand-ins (tests/golden/, produced by oracle/make_golden.py).
"""
from __future__ import annotations

import numpy as np

from . import o3d_standin as o3s

# metric.py:270-281
COLOR_TRANSFORMS = {
    "rgb": np.eye(3),
    "ycc": np.array([[0.2126, 0.7152, 0.0722],
                     [-0.1146, -0.3854, 0.5],
                     [0.5, -0.4542, -0.0458]]),
    "yuv": np.array([[0.25, 0.5, 0.25],
                     [1, 0, -1],
                     [-0.5, 1, -0.5]], dtype=np.float64),
}
# metric.py:293-299
COLOR_PEAK = {"rgb": 255.0, "ycc": 1.0, "yuv": 1.0}


def neighbour_pass(iter_points, search_points):
    """cloud_pair.py:10-42 with n=0: per point of ``iter_points`` the index and the
    SQUARED distance (cloud_pair.py:22-23 returns ``dists[-1]`` unchanged) of its
    nearest neighbour in ``search_points``; canonical tie rule (smallest index)."""
    if len(search_points) == 0:
        raise IndexError("list index out of range")  # idx[-1] on an empty result
    idx, d2 = o3s.exact_knn(search_points, iter_points, 1)
    return idx[:, 0].astype(np.int64), d2[:, 0]


def transform_colors(colors, scheme):
    """metric.py:261-290: rgb is returned untouched, otherwise T @ c per row."""
    if scheme == "rgb":
        return colors
    T = COLOR_TRANSFORMS[scheme]
    c = np.asarray(colors, dtype=np.float64)
    # row-wise matmul(T, c): (T[k,0]*c0 + T[k,1]*c1) + T[k,2]*c2
    return np.stack([(T[k, 0] * c[:, 0] + T[k, 1] * c[:, 1]) + T[k, 2] * c[:, 2] for k in range(3)], axis=1)


class PairOracle:
    """Restates CloudPair.__init__ (cloud_pair.py:54-80) and the getters
    (cloud_pair.py:90-124) for clouds given as numpy arrays."""

    def __init__(self, pts_a, pts_b, col_a=None, col_b=None, nrm_a=None, nrm_b=None, knn=30):
        self.pts = (np.ascontiguousarray(pts_a, dtype=np.float64),
                    np.ascontiguousarray(pts_b, dtype=np.float64))
        self.col = (None if col_a is None else np.asarray(col_a, dtype=np.float64),
                    None if col_b is None else np.asarray(col_b, dtype=np.float64))
        nrm = [nrm_a, nrm_b]
        # cloud_pair.py:61-64 -- normals are estimated for BOTH clouds when missing
        for k in range(2):
            if nrm[k] is None:
                nrm[k] = o3s.estimate_normals_array(self.pts[k], knn)
            else:
                nrm[k] = np.asarray(nrm[k], dtype=np.float64)
        self.nrm = tuple(nrm)
        # cloud_pair.py:67-78
        self.idx = [None, None]
        self.d2 = [None, None]
        self.idx[0], self.d2[0] = neighbour_pass(self.pts[0], self.pts[1])
        self.idx[1], self.d2[1] = neighbour_pass(self.pts[1], self.pts[0])

    # -- getters ------------------------------------------------------------
    def error_vector(self, is_left):
        """cloud_pair.py:90-100 (cloud minus neighbour)."""
        q, s = (0, 1) if is_left else (1, 0)
        return self.pts[q] - self.pts[s][self.idx[q]]

    def neighbour_colors(self, is_left):
        """cloud_pair.py:38-40,120-124."""
        q, s = (0, 1) if is_left else (1, 0)
        return self.col[s][self.idx[q]]

    def boundary_sqrt_distances(self):
        """cloud_pair.py:108-109 + metric.py:182-188: (min, max) of the distance from
        each point of cloud 0 to its nearest OTHER point of cloud 0."""
        pts = self.pts[0]
        if len(pts) < 2:
            d = np.zeros(len(pts))
        else:
            _, d2 = o3s.exact_knn(pts, pts, 2)
            d = np.sqrt(d2[:, 1])
        return (np.min(d), np.max(d))

    def extent(self):
        """cloud_pair.py:111-112."""
        return o3s.minimal_obb_extent(self.pts[0])

    # -- metric.py ----------------------------------------------------------
    def plane_errors(self, is_left):
        """metric.py:124-153, point_to_plane=True.  Quirk Q1: the OTHER cloud's
        normals (metric.py:130) are indexed by the QUERY index i (metric.py:148-152);
        IndexError when the other cloud is shorter."""
        q, s = (0, 1) if is_left else (1, 0)
        E = self.error_vector(is_left)
        nrm = self.nrm[s]
        if len(nrm) < len(E):
            raise IndexError(
                f"index {len(nrm)} is out of bounds for axis 0 with size {len(nrm)}")
        n = nrm[:len(E)]
        return (E[:, 0] * n[:, 0] + E[:, 1] * n[:, 1]) + E[:, 2] * n[:, 2]

    def euclidean_distance(self, is_left, point_to_plane):
        """metric.py:156-179."""
        if not point_to_plane:
            return self.d2[0 if is_left else 1]
        return np.square(self.plane_errors(is_left))

    def geo_mse(self, is_left, point_to_plane):
        """metric.py:213-228."""
        x = self.euclidean_distance(is_left, point_to_plane)
        return np.sum(x, axis=0) / x.shape[0]

    def geo_psnr(self, is_left, point_to_plane, peak=None):
        """metric.py:231-247; peak = max(extent) unless given."""
        if peak is None:
            peak = np.max(self.extent())
        with np.errstate(divide="ignore"):
            return 10 * np.log10(peak ** 2 / self.geo_mse(is_left, point_to_plane))

    def geo_hausdorff(self, is_left, point_to_plane):
        """metric.py:353-366 (max of SQUARED values, quirk Q5)."""
        return np.max(self.euclidean_distance(is_left, point_to_plane), axis=0)

    def geo_hausdorff_psnr(self, is_left, point_to_plane):
        """metric.py:369-386."""
        max_sqrt = self.boundary_sqrt_distances()[1]
        with np.errstate(divide="ignore"):
            return 10 * np.log10(max_sqrt ** 2 / self.geo_hausdorff(is_left, point_to_plane))

    def color_diff(self, is_left, scheme):
        q = 0 if is_left else 1
        o = transform_colors(np.copy(self.col[q]), scheme)
        n = transform_colors(np.copy(self.neighbour_colors(is_left)), scheme)
        return np.subtract(o, n)

    def color_mse(self, is_left, scheme):
        """metric.py:302-333."""
        return np.mean(self.color_diff(is_left, scheme) ** 2, axis=0)

    def color_psnr(self, is_left, scheme):
        """metric.py:336-350."""
        with np.errstate(divide="ignore"):
            return 10 * np.log10(COLOR_PEAK[scheme] ** 2 / self.color_mse(is_left, scheme))

    def color_hausdorff(self, is_left, scheme):
        """metric.py:389-426 (x255 for rgb only, quirk Q6)."""
        diff = self.color_diff(is_left, scheme)
        if scheme == "rgb":
            diff = 255 * diff
        return np.max(diff ** 2, axis=0)

    def color_hausdorff_psnr(self, is_left, scheme):
        """metric.py:429-443."""
        with np.errstate(divide="ignore"):
            return 10 * np.log10(COLOR_PEAK[scheme] ** 2 / self.color_hausdorff(is_left, scheme))


def tie_average(oracle: "PairOracle", is_left, scheme=None):
    """Beyond the reference (MPEG pc_error style): per query, EVERY point of the search cloud at the
    minimal squared distance counts -- its plane error (q - p_t) . n_t with its own normal, squared, and its colour;
    returns (mean over the tied points of the squared plane errors, squared difference between the transformed query
    colour and the transformed MEAN colour of the tied points or None).  Brute force: small clouds only."""
    qi, si = (0, 1) if is_left else (1, 0)
    Q, S = oracle.pts[qi], oracle.pts[si]
    nrm = oracle.nrm[si]
    pe2 = np.empty(len(Q))
    cd2 = None if scheme is None else np.empty((len(Q), 3))
    for i in range(len(Q)):
        d = Q[i] - S
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        tied = np.nonzero(d2 == d2.min())[0]
        e = Q[i] - S[tied]
        n = nrm[tied]
        pe = (e[:, 0] * n[:, 0] + e[:, 1] * n[:, 1]) + e[:, 2] * n[:, 2]
        acc = 0.0
        for v in pe * pe:                     # (index order, one addition at a time: what the kernel's accumulator does)
            acc += v
        pe2[i] = acc * (1.0 / len(tied))
        if scheme is not None:
            csum = np.zeros(3)
            for t in tied:
                csum = csum + oracle.col[si][t]
            cbar = csum * (1.0 / len(tied))
            diff = transform_colors(oracle.col[qi][i:i + 1], scheme)[0] - transform_colors(cbar[None, :], scheme)[0]
            cd2[i] = diff * diff
    return pe2, cd2


def symmetric(lvalue, rvalue, is_proportional):
    """metric.py:475-485."""
    values = [lvalue, rvalue]
    if is_proportional:
        return min(values, key=np.linalg.norm)
    return max(values, key=np.linalg.norm)


def evaluate(oracle: PairOracle, color=None, hausdorff=False, point_to_plane=False, peak=None):
    """options.py:32-174 + calculator.py:97-108: the dict ``as_dict()`` would return,
    keyed by the reference's ``_key()`` tuples."""
    out = {}
    bmin, bmax = oracle.boundary_sqrt_distances()
    out[("MinSqrtDistance",)] = bmin
    out[("MaxSqrtDistance",)] = bmax

    def both(name, fn, prop, *extra):
        lv = fn(True, *extra)
        rv = fn(False, *extra)
        out[(name, True) + extra] = lv
        out[(name, False) + extra] = rv
        out[("SymmetricMetric", name, True) + extra + (name, False) + extra] = symmetric(lv, rv, prop)

    both("GeoMSE", oracle.geo_mse, False, False)
    if peak is None:
        peak = np.max(oracle.extent())
    both("GeoPSNR", lambda l, p: oracle.geo_psnr(l, p, peak), True, False)
    if color is not None:
        both("ColorMSE", oracle.color_mse, False, color)
        both("ColorPSNR", oracle.color_psnr, True, color)
    if point_to_plane:
        both("GeoMSE", oracle.geo_mse, False, True)
        both("GeoPSNR", lambda l, p: oracle.geo_psnr(l, p, peak), True, True)
    if hausdorff:
        both("GeoHausdorffDistance", oracle.geo_hausdorff, False, False)
        both("GeoHausdorffDistancePSNR", oracle.geo_hausdorff_psnr, True, False)
    if hausdorff and point_to_plane:
        both("GeoHausdorffDistance", oracle.geo_hausdorff, False, True)
        both("GeoHausdorffDistancePSNR", oracle.geo_hausdorff_psnr, True, True)
    return out
