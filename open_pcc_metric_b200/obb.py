"""PSNR peak providers.

``minimal_obb_extent`` follows what the reference's peak is made of
(cloud_pair.py:111-112 -> Open3D ``get_minimal_oriented_bounding_box().extent``,
consumed by metric.py:246): Qhull convex hull, then the hull vertices are expressed
in the frame of every hull triangle and the axis-aligned box of least volume wins.
Qhull runs on the host here exactly as it does inside Open3D; the per-facet sweep
(F triangles x V hull vertices) runs on the GPU.  ``aabb_diag`` and ``resolution`` are the two peaks
BASELINE.json's north_star names; they need no hull at all.
"""
from __future__ import annotations

import numpy as np


PREFILTER_MIN_POINTS = 50_000
_DIRS = None


def _directions(n: int = 506) -> np.ndarray:
    """Fixed, deterministic direction set: Fibonacci sphere + the six axis directions."""
    global _DIRS
    if _DIRS is None or len(_DIRS) != n + 6:
        k = np.arange(n) + 0.5
        phi = np.arccos(1 - 2 * k / n)
        th = np.pi * (1 + 5 ** 0.5) * k
        d = np.stack([np.cos(th) * np.sin(phi), np.sin(th) * np.sin(phi), np.cos(phi)], axis=1)
        _DIRS = np.concatenate([d, np.eye(3), -np.eye(3)])
    return _DIRS


def hull_candidates(points: np.ndarray, dev_cloud) -> np.ndarray:
    """Superset of the convex-hull vertices of an indexed device cloud (GPU prefilter): seed points
    = extremes along ~500 directions; everything strictly inside the seeds' hull is dropped."""
    from scipy.spatial import ConvexHull, QhullError
    pts = np.asarray(points, dtype=np.float64)
    seeds = np.unique(dev_cloud.extremes(_directions()))
    try:
        h0 = ConvexHull(pts[seeds])
    except QhullError:          # flat / degenerate seed set: no prefilter
        return pts
    scale = float(np.max(np.abs(pts[seeds]))) + 1.0
    surv, cnt = dev_cloud.outside_hull(h0.equations, eps=1e-9 * scale)
    if cnt != len(surv):
        return pts
    # the kernel appends with an atomic counter: restore a fixed order, Qhull's triangulation of
    # coplanar facets (and with it the box the sweep picks) depends on the input order
    return surv[np.lexsort((surv[:, 2], surv[:, 1], surv[:, 0]))]


def minimal_obb_extent(points: np.ndarray, ctx=None, dev_cloud=None) -> np.ndarray:
    """Extent of the minimal oriented bounding box.  Convex hull: Qhull on the host (exactly what
    Open3D does), fed -- for large clouds whose indexed device copy is given -- only with the
    points the GPU prefilter could not rule out; the F x V facet sweep: ``pccm_obb_sweep`` on the
    GPU (no CPU path).  The prefiltered hull has exactly the vertices of the full hull; Open3D's
    box is the best over the hull TRIANGLES, and how Qhull triangulates coplanar facets depends on
    its input, so extents can differ from the unfiltered run at the 1e-5 level (quirk Q3)."""
    from scipy.spatial import ConvexHull  # Qhull, facets triangulated

    if ctx is None:
        from .geometry import default_context
        ctx = default_context()
    pts = np.ascontiguousarray(points, dtype=np.float64)
    if dev_cloud is not None and len(pts) >= PREFILTER_MIN_POINTS:
        pts = np.ascontiguousarray(hull_candidates(pts, dev_cloud))
    hull = ConvexHull(pts)
    vol, ext = ctx.obb_sweep(pts[hull.vertices], pts[hull.simplices])
    ok = ~np.isnan(vol)
    if not ok.any():
        raise ValueError("degenerate convex hull")
    j = int(np.flatnonzero(ok)[np.argmin(vol[ok])])   # first minimum, like a strict '<' scan over the facets
    return ext[j].copy()


def aabb_diag(aabb_min, aabb_max) -> float:
    d = np.asarray(aabb_max, dtype=np.float64) - np.asarray(aabb_min, dtype=np.float64)
    return float(np.sqrt(np.sum(d * d)))


def resolution_peak(aabb_max, bits: int | None = None) -> float:
    """2^bits - 1; bits inferred from the largest coordinate when not given."""
    if bits is None:
        m = float(np.max(aabb_max))
        bits = max(1, int(np.ceil(np.log2(m + 1)))) if m > 0 else 1
    return float((1 << int(bits)) - 1)
