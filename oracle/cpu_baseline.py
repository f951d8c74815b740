"""CPU timing of the reference path for bench.py -- TEST / MEASUREMENT INFRASTRUCTURE
(see oracle/__init__.py).  The unmodified reference cannot travel to the GPU box
(/root/reference is absent there and Open3D is not installable), so two ports are timed:

``reference_structure``  what the reference really executes: one KD-tree query per point
    from a Python loop (np.apply_along_axis, cloud_pair.py:16-32), a Python ``for`` with
    np.dot per row for D2 (metric.py:146-153) and np.apply_along_axis 3x3 products for the
    colour transform (metric.py:283-290).  Single threaded, as in the reference.
    scipy's cKDTree(leafsize=15) stands in for Open3D's nanoflann tree.
``cpu_best``  the same mathematics batched: cKDTree.query(workers=-1) on all cores and
    whole-array numpy epilogues.  The fair "what a CPU can do" line.
"""
from __future__ import annotations

import os
import time

import numpy as np
from scipy.spatial import cKDTree

from .reference_port import COLOR_TRANSFORMS


def host_cores() -> int:
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def build_trees(A, B):
    """The two KD-trees of a pair (cloud_pair.py:65) and the seconds they took."""
    t0 = time.perf_counter()
    trees = tuple(cKDTree(c.points, leafsize=15) for c in (A, B))
    return trees, time.perf_counter() - t0


def reference_structure(A, B, color_scheme="yuv", point_to_plane=True, sample=50_000, seed=0, prebuilt=None):
    """Times tree builds in full and the per-point work on `sample` queries per direction;
    returns dict(queries_per_s, est_pair_seconds, sample, per_query_us, build_s).
    prebuilt = (trees, seconds) of build_trees: a 10 M-point tree is built once per bench run, not once per arm."""
    rng = np.random.default_rng(seed)
    clouds = (A, B)
    trees, build_s = prebuilt if prebuilt is not None else build_trees(A, B)      # cloud_pair.py:65
    T = COLOR_TRANSFORMS[color_scheme] if color_scheme else None
    n_total = len(A) + len(B)
    work_s = 0.0
    n_sampled = 0
    for q, s in ((0, 1), (1, 0)):
        Q, S = clouds[q], clouds[s]
        m = min(sample, len(Q.points))
        sel = np.sort(rng.choice(len(Q.points), m, replace=False))
        pts = np.ascontiguousarray(Q.points[sel])
        tree = trees[s]
        t0 = time.perf_counter()

        def finder(p):                                                  # cloud_pair.py:16-26
            d, i = tree.query(p.reshape(3), 1)
            return np.array((i, d * d))
        idx, sq = np.apply_along_axis(finder, axis=1, arr=pts).T       # cloud_pair.py:28-32
        idx = idx.astype(int)
        neigh = np.take(S.points, idx, axis=0)
        err = np.subtract(pts, neigh)
        mse = np.sum(sq) / m
        if point_to_plane and S.normals is not None:
            nrm = S.normals
            pe = np.zeros(m)
            for i in range(m):                                          # metric.py:148-152
                pe[i] = np.dot(err[i], nrm[min(sel[i], len(nrm) - 1)])
            mse2 = np.sum(np.square(pe)) / m
        if color_scheme and Q.colors is not None and S.colors is not None:
            def conv(c):
                return np.matmul(T, c)
            oc = np.apply_along_axis(conv, 1, np.copy(Q.colors[sel]))   # metric.py:283-290
            nc = np.apply_along_axis(conv, 1, np.take(S.colors, idx, axis=0))
            cm = np.mean((oc - nc) ** 2, axis=0)
        work_s += time.perf_counter() - t0
        n_sampled += m
    per_query = work_s / n_sampled
    est = build_s + per_query * n_total
    return dict(queries_per_s=n_total / est, est_pair_seconds=est, sample=n_sampled, per_query_us=per_query * 1e6,
                build_s=build_s, measured_s=build_s + work_s, cores=1)


def cpu_best(A, B, color_scheme="yuv", point_to_plane=True, max_queries=None, prebuilt=None):
    """Batched all-core version of the same pair evaluation (full workload unless capped).
    With `prebuilt` = (trees, seconds) and `max_queries` the call is a BOUNDED SAMPLE of a large pair: the first
    max_queries points of each direction are evaluated against the full trees and the pair's time is
    build seconds + sample seconds x (queries of the pair / queries of the sample) -- `sampled` says so."""
    clouds = (A, B)
    if prebuilt is not None and max_queries is not None:
        trees, build_s = prebuilt
        T = COLOR_TRANSFORMS[color_scheme] if color_scheme else None
        est, n_s, t_s = build_s, 0, 0.0
        for q, s in ((0, 1), (1, 0)):
            Q, S = clouds[q], clouds[s]
            m = min(max_queries, len(Q.points))
            t0 = time.perf_counter()
            pts = Q.points[:m]
            _, idx = trees[s].query(pts, 1, workers=-1)
            err = pts - S.points[idx]
            d2 = (err[:, 0] * err[:, 0] + err[:, 1] * err[:, 1]) + err[:, 2] * err[:, 2]
            _ = d2.sum() / len(d2), d2.max()
            if point_to_plane and S.normals is not None:
                nr = S.normals[:m] if len(S.normals) >= m else S.normals[idx]
                _ = np.square((err * nr).sum(1)).sum()
            if color_scheme and Q.colors is not None and S.colors is not None:
                d = Q.colors[:m] @ T.T - S.colors[idx] @ T.T
                _ = (d * d).mean(0)
            dt = time.perf_counter() - t0
            est += dt * (len(Q.points) / m)
            n_s += m
            t_s += dt
        n = len(A.points) + len(B.points)
        return dict(queries_per_s=n / est, seconds=est, queries=n, cores=host_cores(), sampled=True,
                    sample_queries=n_s, sample_seconds=t_s, build_seconds=build_s)
    t0 = time.perf_counter()
    trees = tuple(cKDTree(c.points, leafsize=15) for c in clouds)
    T = COLOR_TRANSFORMS[color_scheme] if color_scheme else None
    n = 0
    for q, s in ((0, 1), (1, 0)):
        Q, S = clouds[q], clouds[s]
        pts = Q.points if max_queries is None else Q.points[:max_queries]
        _, idx = trees[s].query(pts, 1, workers=-1)
        err = pts - S.points[idx]
        d2 = (err[:, 0] * err[:, 0] + err[:, 1] * err[:, 1]) + err[:, 2] * err[:, 2]
        _ = d2.sum() / len(d2), d2.max()
        if point_to_plane and S.normals is not None:
            nr = S.normals[:len(pts)] if len(S.normals) >= len(pts) else S.normals[idx]
            _ = np.square((err * nr).sum(1)).sum()
        if color_scheme and Q.colors is not None and S.colors is not None:
            d = Q.colors[:len(pts)] @ T.T - S.colors[idx] @ T.T
            _ = (d * d).mean(0)
        n += len(pts)
    dt = time.perf_counter() - t0
    return dict(queries_per_s=n / dt, seconds=dt, queries=n, cores=host_cores())
