"""Deterministic synthetic workloads for tests and bench.py.

No dataset can be downloaded in this environment, so every configuration named
in BASELINE.json is generated from a seed: voxelised star-shaped surfaces with
colours and analytic normals (``synth_vox``), a codec-like degradation
(``degrade``) and a LiDAR-style float32 scene (``synth_lidar``).
Base seed 20260101 + configuration index.
"""
from __future__ import annotations

import numpy as np

BASE_SEED = 20260101


class Cloud:
    """Plain container: float64 points (N,3); optional colours in [0,1] (N,3) and
    unit normals (N,3), as Open3D would hand them to the reference."""

    def __init__(self, points, colors=None, normals=None):
        self.points = points
        self.colors = colors
        self.normals = normals

    def __len__(self):
        return len(self.points)


def _directions(rng, n_dir, coef):
    """Unit-radius shape: directions d, radial factor s and its angular derivatives."""
    a, f, g, p, q = coef
    u = rng.random(n_dir)
    v = rng.random(n_dir)
    theta = np.arccos(1.0 - 2.0 * u)  # uniform on the sphere
    phi = 2.0 * np.pi * v
    s = np.zeros(n_dir)
    s_t = np.zeros(n_dir)
    s_p = np.zeros(n_dir)
    for k in range(len(a)):
        st = np.sin(f[k] * theta + p[k])
        sp = np.sin(g[k] * phi + q[k])
        s += a[k] * st * sp
        s_t += a[k] * f[k] * np.cos(f[k] * theta + p[k]) * sp
        s_p += a[k] * g[k] * st * np.cos(g[k] * phi + q[k])
    sin_t, cos_t = np.sin(theta), np.cos(theta)
    sin_p, cos_p = np.sin(phi), np.cos(phi)
    d = np.stack([sin_t * cos_p, sin_t * sin_p, cos_t], axis=1)
    return d, s, s_t, s_p, (sin_t, cos_t, sin_p, cos_p)


def _normals(R, d, s, s_t, s_p, trig):
    sin_t, cos_t, sin_p, cos_p = trig
    n_dir = len(s)
    r = R * (1.0 + 0.25 * s)
    r_t = R * 0.25 * s_t
    r_p = R * 0.25 * s_p
    d_t = np.stack([cos_t * cos_p, cos_t * sin_p, -sin_t], axis=1)
    d_p = np.stack([-sin_t * sin_p, sin_t * cos_p, np.zeros(n_dir)], axis=1)
    S_t = r_t[:, None] * d + r[:, None] * d_t
    S_p = r_p[:, None] * d + r[:, None] * d_p
    nrm = np.cross(S_t, S_p)
    ln = np.linalg.norm(nrm, axis=1)
    bad = ln < 1e-9
    nrm[bad] = d[bad]
    ln[bad] = 1.0
    nrm /= ln[:, None]
    return nrm


def synth_vox(bits: int, target_n: int, seed: int, with_colors=True, with_normals=True,
              oversample: int = 8, phase_shift: float = 0.0, tol: float = 0.02) -> Cloud:
    """Voxelised star-shaped surface with about ``target_n`` occupied voxels."""
    rng0 = np.random.default_rng(np.random.PCG64(seed))
    K = 6
    a = rng0.random(K) / np.arange(1, K + 1)
    f = rng0.integers(1, 9, K).astype(np.float64)
    g = rng0.integers(1, 9, K).astype(np.float64)
    p = rng0.random(K) * 2 * np.pi + phase_shift
    q = rng0.random(K) * 2 * np.pi + phase_shift
    coef = (a, f, g, p, q)
    W = rng0.integers(1, 6, (3, 3)).astype(np.float64)
    phi0 = rng0.random(3) * 2 * np.pi
    size = float(1 << bits)
    centre = size / 2
    R = min(np.sqrt(target_n / (4 * np.pi * 1.15)), size / 2 / 1.45)
    n_dir = oversample * target_n
    rng = np.random.default_rng(np.random.PCG64(seed + 7919))
    d, s, s_t, s_p, trig = _directions(rng, n_dir, coef)  # radius-independent part, once
    for _ in range(8):
        vox = np.floor((R * (1.0 + 0.25 * s))[:, None] * d + centre)
        np.clip(vox, 0, size - 1, out=vox)
        key = (vox[:, 0].astype(np.int64) << 42) | (vox[:, 1].astype(np.int64) << 21) | vox[:, 2].astype(np.int64)
        _, first = np.unique(key, return_index=True)
        n = first.size
        if abs(n - target_n) <= tol * target_n or R >= size / 2 / 1.45:
            break
        R = min(R * np.sqrt(target_n / n), size / 2 / 1.45)
    pts = vox[first]
    nrm = _normals(R, d[first], s[first], s_t[first], s_p[first], tuple(t[first] for t in trig))
    perm = rng.permutation(len(pts))  # file order must not be sorted order
    pts = np.ascontiguousarray(pts[perm])
    nrm = np.ascontiguousarray(nrm[perm])
    cols = None
    if with_colors:
        arg = (pts / size * 2 * np.pi) @ W.T + phi0
        c = 128 + 90 * np.sin(arg) + rng.normal(0, 6, pts.shape)
        cols = np.clip(c, 0, 255).astype(np.uint8).astype(np.float64) / 255.0
    return Cloud(pts, cols, nrm if with_normals else None)


def degrade(cloud: Cloud, step: int, seed: int, bits: int | None = None, dedup: bool = True) -> Cloud:
    """Codec-like degradation: coarser lattice + sparse jitter.  ``dedup=True`` removes
    coinciding points and shuffles (a decoded cloud); ``dedup=False`` keeps one output point
    per input point in the same order (the shape of the reference's own smoke test,
    tests/unit/test_metric.py:201-223, and the only shape for which the reference's D2 is
    defined in both directions -- quirk Q1)."""
    rng = np.random.default_rng(np.random.PCG64(seed + 104729))
    A = cloud.points
    n = len(A)
    B = np.round(A / step) * step
    jit = rng.integers(-1, 2, (n, 3)).astype(np.float64)
    mask = rng.random(n) < 0.25
    B[mask] += jit[mask]
    hi = float((1 << bits) - 1) if bits is not None else max(float(A.max()), 0.0)
    np.clip(B, 0, hi, out=B)
    if dedup:
        key = (B[:, 0].astype(np.int64) << 42) | (B[:, 1].astype(np.int64) << 21) | B[:, 2].astype(np.int64)
        _, first = np.unique(key, return_index=True)
        first = first[rng.permutation(first.size)]
    else:
        first = np.arange(n)
    pts = np.ascontiguousarray(B[first])
    cols = None
    if cloud.colors is not None:
        c8 = np.rint(cloud.colors[first] * 255.0) + rng.normal(0, 4, (first.size, 3))
        cols = np.clip(np.rint(c8), 0, 255).astype(np.uint8).astype(np.float64) / 255.0
    nrm = None
    if cloud.normals is not None:
        nrm = np.ascontiguousarray(cloud.normals[first])
    return Cloud(pts, cols, nrm)


def synth_pair(bits: int, target_n: int, seed: int, step: int = 2, with_colors=True,
               with_normals=True, phase_shift: float = 0.0, dedup: bool = True, oversample: int = 8):
    A = synth_vox(bits, target_n, seed, with_colors, with_normals, phase_shift=phase_shift, oversample=oversample)
    B = degrade(A, step, seed, bits, dedup=dedup)
    return A, B


def synth_lidar(target_n: int, seed: int, dedup: bool = True):
    """LiDAR-style float32 scene (config 5): rolling ground plane + boxes + cylinders,
    density ~ 1/max(r,2)^2 from the origin.  Returns (A, B) float32-valued clouds
    (stored as float32 arrays) where B is A snapped to a 2 cm lattice, de-duplicated
    and jittered by N(0, 5 mm).  dedup=False keeps one (separately jittered) point of B per
    point of A, so that both clouds have target_n points."""
    rng = np.random.default_rng(np.random.PCG64(seed))
    n_ground = int(0.6 * target_n)
    n_obj = target_n - n_ground
    # ground: log-uniform radius gives areal density ~ 1/r^2
    r = np.exp(rng.uniform(np.log(0.5), np.log(141.0), n_ground))
    ang = rng.uniform(0, 2 * np.pi, n_ground)
    x = np.clip(r * np.cos(ang), -100, 100)
    y = np.clip(r * np.sin(ang), -100, 100)
    z = 0.05 * np.sin(x / 7.0) + rng.normal(0, 0.02, n_ground)
    parts = [np.stack([x, y, z], axis=1)]
    n_box, n_cyl = 200, 100
    per = max(1, n_obj // (n_box + n_cyl))
    cx = rng.uniform(-90, 90, n_box + n_cyl)
    cy = rng.uniform(-90, 90, n_box + n_cyl)
    for i in range(n_box):
        sx, sy, sz = rng.uniform(1, 8), rng.uniform(1, 8), rng.uniform(1.5, 10)
        face = rng.integers(0, 4, per)
        u = rng.random(per)
        v = rng.random(per)
        px = np.where(face < 2, (face * 1.0) * sx, u * sx)
        py = np.where(face < 2, u * sy, (face - 2.0) * sy)
        parts.append(np.stack([cx[i] + px - sx / 2, cy[i] + py - sy / 2, v * sz], axis=1))
    for i in range(n_cyl):
        rad, hgt = rng.uniform(0.15, 0.6), rng.uniform(3, 12)
        t = rng.uniform(0, 2 * np.pi, per)
        parts.append(np.stack([cx[n_box + i] + rad * np.cos(t), cy[n_box + i] + rad * np.sin(t),
                               rng.random(per) * hgt], axis=1))
    A = np.concatenate(parts).astype(np.float32)
    A = A[rng.permutation(len(A))]
    q = np.round(A.astype(np.float64) / 0.02)
    if dedup:
        key = ((q[:, 0] + 8192).astype(np.int64) << 42) | ((q[:, 1] + 8192).astype(np.int64) << 21) | (q[:, 2] + 8192).astype(np.int64)
        _, first = np.unique(key, return_index=True)
        first = first[rng.permutation(first.size)]
    else:
        first = rng.permutation(len(A))
    B = (q[first] * 0.02 + rng.normal(0, 0.005, (first.size, 3))).astype(np.float32)
    return Cloud(np.ascontiguousarray(A)), Cloud(np.ascontiguousarray(B))
