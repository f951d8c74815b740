"""Point-cloud container with the slice of Open3D's ``o3d.geometry.PointCloud``
surface that the reference touches (cloud_pair.py:31-40,61-64,109,112;
metric.py:95,98; handler.py:57), backed by libpccm.so instead of Open3D.

The geometry methods run on the GPU (k-NN + PCA normals, self nearest-neighbour
distances); only the convex hull of ``get_minimal_oriented_bounding_box`` runs on
the host (Qhull via scipy, as Open3D itself does it with Qhull on the CPU).
"""
from __future__ import annotations

import numpy as np

from . import _native as N
from . import obb as _obb


class Vector3dVector(np.ndarray):
    """float64 (N, 3) array; constructing one copies, like the pybind type."""

    def __new__(cls, arr=None):
        a = np.zeros((0, 3)) if arr is None else np.array(arr, dtype=np.float64, copy=True)
        return a.reshape(-1, 3).view(cls)


_CTX = {}


def default_context(device: int | None = None) -> N.Context:
    """Process-wide pccm context per device (rank-local GPU under torchrun)."""
    if device is None:
        import os
        device = int(os.environ.get("PCCM_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    if device not in _CTX:
        _CTX[device] = N.Context(device)
    return _CTX[device]


class OrientedBoundingBox:
    def __init__(self, extent):
        self.extent = extent


class PointCloud:
    """points / colors / normals as float64 (N, 3); colours in [0, 1]."""

    def __init__(self, points=None, colors=None, normals=None):
        self._points = Vector3dVector(points)
        self._colors = Vector3dVector(colors)
        self._normals = Vector3dVector(normals)

    # attribute surface ------------------------------------------------------
    @property
    def points(self):
        return self._points

    @points.setter
    def points(self, v):
        self._points = Vector3dVector(v)

    @property
    def colors(self):
        return self._colors

    @colors.setter
    def colors(self, v):
        self._colors = Vector3dVector(v)

    @property
    def normals(self):
        return self._normals

    @normals.setter
    def normals(self, v):
        self._normals = Vector3dVector(v)

    def __len__(self):
        return len(self._points)

    def has_points(self):
        return len(self._points) > 0

    def has_colors(self):
        return len(self._points) > 0 and len(self._colors) == len(self._points)

    def has_normals(self):
        return len(self._points) > 0 and len(self._normals) == len(self._points)

    # GPU-backed geometry ----------------------------------------------------
    def _indexed(self, ctx=None) -> N.Cloud:
        ctx = ctx or default_context()
        c = ctx.cloud(np.asarray(self._points))
        c.build_index()
        return c

    def estimate_normals(self, knn: int = 30, ctx=None):
        """Open3D defaults: KDTreeSearchParamKNN(knn=30), fast_normal_computation=True,
        no orientation step (reference call: cloud_pair.py:62,64)."""
        if len(self._points) == 0:
            return
        c = self._indexed(ctx)
        try:
            c.estimate_normals(knn)
            self._normals = Vector3dVector(c.get_normals())
        finally:
            c.close()

    def compute_nearest_neighbor_distance(self, ctx=None):
        """Distance of every point to its nearest other point (cloud_pair.py:109)."""
        n = len(self._points)
        if n < 2:
            return np.zeros((n,), dtype=np.float64)
        c = self._indexed(ctx)
        try:
            return c.self_nn_minmax(per_point=True)[2]
        finally:
            c.close()

    def get_minimal_oriented_bounding_box(self, robust: bool = False):
        """cloud_pair.py:112 -- only ``.extent`` is consumed."""
        return OrientedBoundingBox(_obb.minimal_obb_extent(np.asarray(self._points), default_context()))

    def get_axis_aligned_bounding_box_extent(self):
        p = np.asarray(self._points)
        return p.max(axis=0) - p.min(axis=0)
