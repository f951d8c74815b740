"""PSNR peak providers.

``minimal_obb_extent`` follows what the reference's peak is made of
(cloud_pair.py:111-112 -> Open3D ``get_minimal_oriented_bounding_box().extent``,
consumed by metric.py:246): Qhull convex hull, then the hull vertices are expressed
in the frame of every hull triangle and the axis-aligned box of least volume wins.
Qhull runs on the host here exactly as it does inside Open3D; the per-facet sweep is
a batched matrix product.  ``aabb_diag`` and ``resolution`` are the two peaks
BASELINE.json's north_star names; they need no hull at all.
"""
from __future__ import annotations

import numpy as np


def minimal_obb_extent(points: np.ndarray, facet_chunk: int = 256) -> np.ndarray:
    from scipy.spatial import ConvexHull  # Qhull, facets triangulated

    pts = np.ascontiguousarray(points, dtype=np.float64)
    hull = ConvexHull(pts)
    hv = pts[hull.vertices]                       # (V, 3)
    tri = pts[hull.simplices]                     # (F, 3, 3)
    a = tri[:, 0]
    e1 = tri[:, 1] - a
    e2 = tri[:, 2] - a
    w = np.cross(e1, e2)
    v = np.cross(w, e1)
    with np.errstate(invalid="ignore", divide="ignore"):
        frames = np.stack([e1 / np.linalg.norm(e1, axis=1, keepdims=True),
                           v / np.linalg.norm(v, axis=1, keepdims=True),
                           w / np.linalg.norm(w, axis=1, keepdims=True)], axis=2)  # columns u, v, w
    best_vol = None
    best_ext = None
    for s in range(0, len(frames), facet_chunk):
        fr = frames[s:s + facet_chunk]
        inv = np.linalg.inv(fr)                                   # (f, 3, 3)
        rel = hv[None, :, :] - a[s:s + facet_chunk, None, :]      # (f, V, 3)
        loc = np.einsum("fij,fvj->fvi", inv, rel) + a[s:s + facet_chunk, None, :]
        ext = loc.max(axis=1) - loc.min(axis=1)                   # (f, 3)
        vol = ext[:, 0] * ext[:, 1] * ext[:, 2]
        j = int(np.argmin(vol))                                   # first minimum, like a strict '<' scan
        if best_vol is None or vol[j] < best_vol:
            best_vol = vol[j]
            best_ext = ext[j].copy()
    return best_ext


def aabb_diag(aabb_min, aabb_max) -> float:
    d = np.asarray(aabb_max, dtype=np.float64) - np.asarray(aabb_min, dtype=np.float64)
    return float(np.sqrt(np.sum(d * d)))


def resolution_peak(aabb_max, bits: int | None = None) -> float:
    """2^bits - 1; bits inferred from the largest coordinate when not given."""
    if bits is None:
        m = float(np.max(aabb_max))
        bits = max(1, int(np.ceil(np.log2(m + 1)))) if m > 0 else 1
    return float((1 << int(bits)) - 1)
