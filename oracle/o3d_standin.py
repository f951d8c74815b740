"""Open3D 0.18.0 stand-in (TEST INFRASTRUCTURE -- see oracle/__init__.py).

The reference calls exactly these Open3D entry points:

  o3d.geometry.PointCloud()                         cloud_pair.py:35
  .points / .colors / .normals                      cloud_pair.py:31,34,39,92-99,115-124; metric.py:95,98
  .has_normals() / .has_colors()                    cloud_pair.py:61,63,38
  .estimate_normals()                               cloud_pair.py:62,64
  .compute_nearest_neighbor_distance()              cloud_pair.py:109
  .get_minimal_oriented_bounding_box().extent       cloud_pair.py:112
  o3d.geometry.KDTreeFlann(cloud)                   cloud_pair.py:65
  KDTreeFlann.search_knn_vector_3d(p, k)            cloud_pair.py:22
  o3d.utility.Vector3dVector(ndarray)               cloud_pair.py:36,40

Open3D 0.18.0 (pinned by the reference's poetry.lock:1131-1132) is a third-party
wheel that is absent from /root/reference and cannot be installed here, so each
entry point is restated from Open3D's published algorithm:

* k-NN: exact Euclidean k nearest neighbours in float64.  nanoflann's L2 adaptor
  accumulates ``diff*diff`` per dimension in order, i.e. ``(dx*dx + dy*dy) + dz*dz``.
  Results are sorted by distance.  TIE RULE (documented deviation): nanoflann keeps
  the first-visited point among equal distances, which depends on its tree layout;
  the oracle and the GPU path both use the canonical rule "smaller original index
  wins among equal squared distances" (sort key ``(d2, index)``).
* estimate_normals(): defaults ``KDTreeSearchParamKNN(knn=30)``,
  ``fast_normal_computation=True``: covariance of the k-NN set (self included) from
  nine running cumulants divided by k (``utility::ComputeCovariance``), then the
  analytic robust symmetric 3x3 eigen solver (``FastEigen3x3``, Eberly) for the
  eigenvector of the smallest eigenvalue; identity covariance when fewer than three
  neighbours; (0, 0, 1) when the solver returns the zero vector; no orientation.
* compute_nearest_neighbor_distance(): sqrt of the second entry of the 2-NN result
  in the cloud's own tree; zeros when the cloud has fewer than two points.
* get_minimal_oriented_bounding_box(): Qhull convex hull (scipy.spatial.ConvexHull
  is Qhull), then for every hull triangle the hull vertices are rotated into the
  triangle's frame and the axis-aligned box of smallest volume wins
  (``OrientedBoundingBox::CreateFromPointsMinimal``).
"""
from __future__ import annotations

import math
import sys
import types

import numpy as np

try:  # scipy is only used to accelerate candidate generation and for Qhull
    from scipy.spatial import ConvexHull, cKDTree
except Exception:  # pragma: no cover
    ConvexHull = None
    cKDTree = None


# --------------------------------------------------------------------------
# exact canonical k-NN
# --------------------------------------------------------------------------
def sq_dists(q: np.ndarray, p: np.ndarray) -> np.ndarray:
    """(Q,3),(N,3) -> (Q,N) squared distances, nanoflann L2 order."""
    dx = q[:, None, 0] - p[None, :, 0]
    dy = q[:, None, 1] - p[None, :, 1]
    dz = q[:, None, 2] - p[None, :, 2]
    return (dx * dx + dy * dy) + dz * dz


def _brute_knn(points, queries, k):
    n = points.shape[0]
    k = min(k, n)
    nq = queries.shape[0]
    idx = np.empty((nq, k), dtype=np.int64)
    d2 = np.empty((nq, k), dtype=np.float64)
    chunk = max(1, int(2e7) // max(n, 1))
    for s in range(0, nq, chunk):
        dd = sq_dists(queries[s:s + chunk], points)
        if k == 1:
            j = dd.argmin(axis=1)[:, None]  # first minimum == smallest index
        else:
            j = np.argsort(dd, axis=1, kind="stable")[:, :k]
        idx[s:s + chunk] = j
        d2[s:s + chunk] = np.take_along_axis(dd, j, axis=1)
    return idx, d2


def exact_knn(points: np.ndarray, queries: np.ndarray, k: int, tree=None):
    """Canonical exact k-NN: rows sorted by (d2, original index).

    Returns (idx[Q, k'], d2[Q, k']) with k' = min(k, N).
    Small problems are brute force.  Large ones use cKDTree only to propose
    candidates; squared distances are always recomputed from coordinates and any
    query whose candidate list cannot prove completeness falls back to brute force.
    """
    points = np.ascontiguousarray(points, dtype=np.float64)
    queries = np.ascontiguousarray(queries, dtype=np.float64)
    n = points.shape[0]
    nq = queries.shape[0]
    k = min(k, n)
    if k == 0 or nq == 0:
        return (np.empty((nq, 0), dtype=np.int64), np.empty((nq, 0), dtype=np.float64))
    if cKDTree is None or n <= 2048 or n * nq <= int(4e7):
        return _brute_knn(points, queries, k)
    if tree is None:
        tree = cKDTree(points, leafsize=15)
    extra = 10
    while True:
        kk = min(n, k + extra)
        _, cand = tree.query(queries, k=kk, workers=-1)
        cand = cand.reshape(nq, kk).astype(np.int64)
        cp = points[cand]  # (Q, kk, 3)
        dx = queries[:, None, 0] - cp[:, :, 0]
        dy = queries[:, None, 1] - cp[:, :, 1]
        dz = queries[:, None, 2] - cp[:, :, 2]
        dd = (dx * dx + dy * dy) + dz * dz
        order = np.lexsort((cand, dd), axis=1)
        cand = np.take_along_axis(cand, order, axis=1)
        dd = np.take_along_axis(dd, order, axis=1)
        if kk == n:
            return cand[:, :k].copy(), dd[:, :k].copy()
        # complete iff the farthest candidate is strictly farther than the k-th
        # (margin covers cKDTree's own rounding when it ranked the candidates)
        ok = dd[:, -1] > dd[:, k - 1] * (1 + 1e-9) + 1e-300
        bad = np.nonzero(~ok)[0]
        if bad.size > max(64, nq // 50) and kk < n:
            extra *= 4
            continue
        idx = cand[:, :k].copy()
        d2 = dd[:, :k].copy()
        if bad.size:
            bi, bd = _brute_knn(points, queries[bad], k)
            idx[bad] = bi
            d2[bad] = bd
        return idx, d2


# --------------------------------------------------------------------------
# Open3D EstimateNormals pieces
# --------------------------------------------------------------------------
def compute_covariance(points: np.ndarray, indices) -> np.ndarray:
    """utility::ComputeCovariance: nine cumulants, divide by count, E[xx^T]-E[x]E[x]^T."""
    if len(indices) == 0:
        return np.eye(3)
    c = [0.0] * 9
    for idx in indices:
        x, y, z = (float(v) for v in points[idx])
        c[0] += x
        c[1] += y
        c[2] += z
        c[3] += x * x
        c[4] += x * y
        c[5] += x * z
        c[6] += y * y
        c[7] += y * z
        c[8] += z * z
    cnt = float(len(indices))
    c = [v / cnt for v in c]
    cov = np.empty((3, 3))
    cov[0, 0] = c[3] - c[0] * c[0]
    cov[1, 1] = c[6] - c[1] * c[1]
    cov[2, 2] = c[8] - c[2] * c[2]
    cov[0, 1] = cov[1, 0] = c[4] - c[0] * c[1]
    cov[0, 2] = cov[2, 0] = c[5] - c[0] * c[2]
    cov[1, 2] = cov[2, 1] = c[7] - c[1] * c[2]
    return cov


def _cross(a, b):
    return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])


def _dot(a, b):
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]


def _eigenvector0(A, eval0):
    row0 = (A[0][0] - eval0, A[0][1], A[0][2])
    row1 = (A[0][1], A[1][1] - eval0, A[1][2])
    row2 = (A[0][2], A[1][2], A[2][2] - eval0)
    r0xr1 = _cross(row0, row1)
    r0xr2 = _cross(row0, row2)
    r1xr2 = _cross(row1, row2)
    d0 = _dot(r0xr1, r0xr1)
    d1 = _dot(r0xr2, r0xr2)
    d2 = _dot(r1xr2, r1xr2)
    dmax = d0
    imax = 0
    if d1 > dmax:
        dmax = d1
        imax = 1
    if d2 > dmax:
        imax = 2
    with np.errstate(all="ignore"):
        if imax == 0:
            s = math.sqrt(d0)
            return tuple(np.float64(v) / np.float64(s) for v in r0xr1)
        if imax == 1:
            s = math.sqrt(d1)
            return tuple(np.float64(v) / np.float64(s) for v in r0xr2)
        s = math.sqrt(d2)
        return tuple(np.float64(v) / np.float64(s) for v in r1xr2)


def _eigenvector1(A, evec0, eval1):
    with np.errstate(all="ignore"):
        if abs(evec0[0]) > abs(evec0[1]):
            inv_length = np.float64(1) / np.sqrt(np.float64(evec0[0] * evec0[0] + evec0[2] * evec0[2]))
            U = (-evec0[2] * inv_length, 0.0, evec0[0] * inv_length)
        else:
            inv_length = np.float64(1) / np.sqrt(np.float64(evec0[1] * evec0[1] + evec0[2] * evec0[2]))
            U = (0.0, evec0[2] * inv_length, -evec0[1] * inv_length)
        V = _cross(evec0, U)
        AU = (A[0][0] * U[0] + A[0][1] * U[1] + A[0][2] * U[2],
              A[0][1] * U[0] + A[1][1] * U[1] + A[1][2] * U[2],
              A[0][2] * U[0] + A[1][2] * U[1] + A[2][2] * U[2])
        AV = (A[0][0] * V[0] + A[0][1] * V[1] + A[0][2] * V[2],
              A[0][1] * V[0] + A[1][1] * V[1] + A[1][2] * V[2],
              A[0][2] * V[0] + A[1][2] * V[1] + A[2][2] * V[2])
        m00 = U[0] * AU[0] + U[1] * AU[1] + U[2] * AU[2] - eval1
        m01 = U[0] * AV[0] + U[1] * AV[1] + U[2] * AV[2]
        m11 = V[0] * AV[0] + V[1] * AV[1] + V[2] * AV[2] - eval1
        a00, a01, a11 = abs(m00), abs(m01), abs(m11)
        if a00 >= a11:
            if max(a00, a01) > 0:
                if a00 >= a01:
                    m01 = m01 / m00
                    m00 = 1 / math.sqrt(1 + m01 * m01)
                    m01 = m01 * m00
                else:
                    m00 = m00 / m01
                    m01 = 1 / math.sqrt(1 + m00 * m00)
                    m00 = m00 * m01
                return tuple(m01 * U[i] - m00 * V[i] for i in range(3))
            return U
        if max(a11, a01) > 0:
            if a11 >= a01:
                m01 = m01 / m11
                m11 = 1 / math.sqrt(1 + m01 * m01)
                m01 = m01 * m11
            else:
                m11 = m11 / m01
                m01 = 1 / math.sqrt(1 + m11 * m11)
                m11 = m11 * m01
            return tuple(m11 * U[i] - m01 * V[i] for i in range(3))
        return U


def fast_eigen_3x3(cov: np.ndarray):
    """geometry::FastEigen3x3 (EstimateNormals.cpp): eigenvector of the smallest
    eigenvalue of a symmetric 3x3, Eberly's robust analytic method."""
    max_coeff = float(np.max(cov))
    if max_coeff == 0:
        return (0.0, 0.0, 0.0)
    A = [[float(cov[i, j]) / max_coeff for j in range(3)] for i in range(3)]
    norm = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2]
    if norm > 0:
        q = (A[0][0] + A[1][1] + A[2][2]) / 3
        b00 = A[0][0] - q
        b11 = A[1][1] - q
        b22 = A[2][2] - q
        p = math.sqrt((b00 * b00 + b11 * b11 + b22 * b22 + norm * 2) / 6)
        c00 = b11 * b22 - A[1][2] * A[1][2]
        c01 = A[0][1] * b22 - A[1][2] * A[0][2]
        c02 = A[0][1] * A[1][2] - b11 * A[0][2]
        det = (b00 * c00 - A[0][1] * c01 + A[0][2] * c02) / (p * p * p)
        half_det = det * 0.5
        half_det = min(max(half_det, -1.0), 1.0)
        angle = math.acos(half_det) / 3.0
        two_thirds_pi = 2.09439510239319549
        beta2 = math.cos(angle) * 2
        beta0 = math.cos(angle + two_thirds_pi) * 2
        beta1 = -(beta0 + beta2)
        e0 = q + p * beta0
        e1 = q + p * beta1
        e2 = q + p * beta2
        if half_det >= 0:
            evec2 = _eigenvector0(A, e2)
            if e2 < e0 and e2 < e1:
                return evec2
            evec1 = _eigenvector1(A, evec2, e1)
            if e1 < e0 and e1 < e2:
                return evec1
            return _cross(evec1, evec2)
        evec0 = _eigenvector0(A, e0)
        if e0 < e1 and e0 < e2:
            return evec0
        evec1 = _eigenvector1(A, evec0, e1)
        if e1 < e0 and e1 < e2:
            return evec1
        return _cross(evec0, evec1)
    # diagonal matrix (A *= max_coeff restores the input; only comparisons follow)
    a00, a11, a22 = float(cov[0, 0]), float(cov[1, 1]), float(cov[2, 2])
    # Open3D compares the rescaled A; (x / m) * m may differ from x in the last
    # bit, so rescale exactly as it does.
    a00, a11, a22 = (a00 / max_coeff) * max_coeff, (a11 / max_coeff) * max_coeff, (a22 / max_coeff) * max_coeff
    if a00 < a11 and a00 < a22:
        return (1.0, 0.0, 0.0)
    if a11 < a00 and a11 < a22:
        return (0.0, 1.0, 0.0)
    return (0.0, 0.0, 1.0)


def estimate_normals_array(points: np.ndarray, knn: int = 30, nn_idx=None) -> np.ndarray:
    """PointCloud::EstimateNormals(KDTreeSearchParamKNN(knn), fast=True) for a cloud
    that has no normals yet.  ``nn_idx`` (N, k') may be supplied (canonical k-NN)."""
    points = np.ascontiguousarray(points, dtype=np.float64)
    n = points.shape[0]
    normals = np.empty((n, 3), dtype=np.float64)
    if n == 0:
        return normals
    if nn_idx is None:
        nn_idx, _ = exact_knn(points, points, knn)
    for i in range(n):
        ind = nn_idx[i]
        if len(ind) >= 3:
            cov = compute_covariance(points, ind)
        else:
            cov = np.eye(3)
        nrm = fast_eigen_3x3(cov)
        nv = np.array(nrm, dtype=np.float64)
        if not (np.linalg.norm(nv) != 0.0):  # zero vector (NaN keeps NaN like Eigen)
            if np.linalg.norm(nv) == 0.0:
                nv = np.array([0.0, 0.0, 1.0])
        normals[i] = nv
    return normals


# --------------------------------------------------------------------------
# minimal oriented bounding box
# --------------------------------------------------------------------------
def minimal_obb_extent(points: np.ndarray) -> np.ndarray:
    """OrientedBoundingBox::CreateFromPointsMinimal(...).extent_."""
    points = np.ascontiguousarray(points, dtype=np.float64)
    hull = ConvexHull(points)  # Qhull, facets triangulated ("Qt")
    verts_idx = hull.vertices
    verts = points[verts_idx]
    min_vol = -1.0
    best_extent = None
    for tri in hull.simplices:
        a, b, c = points[tri[0]], points[tri[1]], points[tri[2]]
        u = b - a
        v = c - a
        w = np.cross(u, v)
        v = np.cross(w, u)
        u = u / np.linalg.norm(u)
        v = v / np.linalg.norm(v)
        w = w / np.linalg.norm(w)
        m_rot = np.stack([u, v, w], axis=1)  # columns u, v, w
        local = (np.linalg.inv(m_rot) @ (verts - a).T).T + a
        ext = local.max(axis=0) - local.min(axis=0)
        vol = ext[0] * ext[1] * ext[2]
        if min_vol == -1.0 or vol < min_vol:
            min_vol = vol
            best_extent = ext
    return best_extent


# --------------------------------------------------------------------------
# the fake module surface
# --------------------------------------------------------------------------
class Vector3dVector(np.ndarray):
    """numpy-backed stand-in for o3d.utility.Vector3dVector (copies like pybind)."""

    def __new__(cls, arr=None):
        if arr is None:
            arr = np.zeros((0, 3))
        a = np.array(arr, dtype=np.float64, copy=True)
        if a.ndim != 2 or a.shape[1] != 3:
            a = a.reshape(-1, 3)
        return a.view(cls)


class _OBB:
    def __init__(self, extent):
        self.extent = extent


class PointCloud:
    def __init__(self, points=None):
        self._points = Vector3dVector(points)
        self._colors = Vector3dVector()
        self._normals = Vector3dVector()

    points = property(lambda self: self._points,
                      lambda self, v: setattr(self, "_points", Vector3dVector(v)))
    colors = property(lambda self: self._colors,
                      lambda self, v: setattr(self, "_colors", Vector3dVector(v)))
    normals = property(lambda self: self._normals,
                       lambda self, v: setattr(self, "_normals", Vector3dVector(v)))

    def has_points(self):
        return len(self._points) > 0

    def has_colors(self):
        return len(self._points) > 0 and len(self._colors) == len(self._points)

    def has_normals(self):
        return len(self._points) > 0 and len(self._normals) == len(self._points)

    def estimate_normals(self, knn: int = 30):
        self._normals = Vector3dVector(estimate_normals_array(np.asarray(self._points), knn))

    def compute_nearest_neighbor_distance(self):
        pts = np.asarray(self._points)
        n = len(pts)
        if n < 2:
            return np.zeros((n,), dtype=np.float64)
        _, d2 = exact_knn(pts, pts, 2)
        return np.sqrt(d2[:, 1])

    def get_minimal_oriented_bounding_box(self, robust: bool = False):
        return _OBB(minimal_obb_extent(np.asarray(self._points)))


class KDTreeFlann:
    def __init__(self, cloud: PointCloud):
        self._pts = np.ascontiguousarray(np.asarray(cloud.points), dtype=np.float64)
        self._tree = None
        if cKDTree is not None and len(self._pts) > 2048:
            self._tree = cKDTree(self._pts, leafsize=15)
        self._cache = None

    def search_knn_vector_3d(self, query, knn):
        q = np.asarray(query, dtype=np.float64).reshape(1, 3)
        n = self._pts.shape[0]
        if n == 0:
            return [0, [], []]
        if self._tree is None:
            idx, d2 = _brute_knn(self._pts, q, knn)
        else:
            idx, d2 = _tree_knn_single(self._tree, self._pts, q, knn)
        return [idx.shape[1], [int(i) for i in idx[0]], [float(d) for d in d2[0]]]


def _tree_knn_single(tree, pts, q, k):
    n = pts.shape[0]
    k = min(k, n)
    extra = 10
    while True:
        kk = min(n, k + extra)
        _, cand = tree.query(q[0], k=kk)
        cand = np.atleast_1d(cand).astype(np.int64)
        cp = pts[cand]
        dx = q[0, 0] - cp[:, 0]
        dy = q[0, 1] - cp[:, 1]
        dz = q[0, 2] - cp[:, 2]
        dd = (dx * dx + dy * dy) + dz * dz
        order = np.lexsort((cand, dd))
        cand, dd = cand[order], dd[order]
        if kk == n or dd[-1] > dd[k - 1] * (1 + 1e-9) + 1e-300:
            return cand[None, :k], dd[None, :k]
        extra *= 4


def install_fake_open3d():
    """Register this stand-in as ``open3d`` in sys.modules so that the UNMODIFIED
    reference package (which does ``import open3d as o3d``; cloud_pair.py:3,
    metric.py:4) can be imported in the build container."""
    mod = types.ModuleType("open3d")
    geometry = types.ModuleType("open3d.geometry")
    utility = types.ModuleType("open3d.utility")
    geometry.PointCloud = PointCloud
    geometry.KDTreeFlann = KDTreeFlann
    utility.Vector3dVector = Vector3dVector
    mod.geometry = geometry
    mod.utility = utility
    mod.__version__ = "0.18.0-standin"
    sys.modules["open3d"] = mod
    sys.modules["open3d.geometry"] = geometry
    sys.modules["open3d.utility"] = utility
    return mod
