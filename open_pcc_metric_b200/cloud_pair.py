"""``CloudPair`` -- drop-in for open_pcc_metric.cloud_pair.CloudPair
(reference cloud_pair.py:45-124) whose geometry runs on a B200 through libpccm.so.

Same constructor, same public attributes (``clouds``, ``origin_cloud``,
``reconst_cloud``) and the same twelve getters.  What the reference does with one
Open3D call per point (cloud_pair.py:16-32) happens here as: upload both clouds,
build a pencil-grid index per cloud, and run one fused kernel per direction that
finds every nearest neighbour and reduces the D1 / D2 / colour sums and maxima on
chip (``fused``).  Per-point arrays (indices, squared distances, error vectors,
neighbour colours) are produced only when a getter asks for them.

Extra keyword-only options (all default to the reference's behaviour):
  normals_mode  "reference" -> D2 uses the other cloud's normal at the QUERY index
                (metric.py:130,148-152, quirk Q1); "neighbour" -> normal of the match.
  ties          "first" (the reference keeps ONE nearest neighbour, cloud_pair.py:22-23; here: the one with the smallest
                index) | "average" -> plane error and colour are averaged over every point at the minimal distance,
                each tied point with its own normal (MPEG pc_error style).  Affects the fused D2 /
                colour reductions only (GeoMSE / GeoPSNR point-to-plane, ColorMSE / ColorPSNR, their Hausdorff forms).
  peak          "obb" (reference, cloud_pair.py:111-112) | "aabb_diag" | "resolution"
  eager_normals estimate missing normals in the constructor like cloud_pair.py:61-64
                (default: on first use; values are identical).
  rank, world, group   multi-GPU: this process reduces slice ``rank`` of ``world``
                of every query cloud and the partial sums are exchanged with
                torch.distributed (NCCL) -- see SURVEY.md section 8(e).
"""
from __future__ import annotations

import typing

import numpy as np

from . import _native as N
from . import obb as _obb
from .geometry import default_context

_COLOR_TRANSFORMS = {
    # metric.py:270-281; "rgb" is the identity (metric.py:266-267 returns the input)
    "rgb": ((1.0, 0.0, 0.0), (0.0, 1.0, 0.0), (0.0, 0.0, 1.0)),
    "ycc": ((0.2126, 0.7152, 0.0722), (-0.1146, -0.3854, 0.5), (0.5, -0.4542, -0.0458)),
    "yuv": ((0.25, 0.5, 0.25), (1.0, 0.0, -1.0), (-0.5, 1.0, -0.5)),
}


def _attr(cloud, name):
    v = getattr(cloud, name, None)
    if v is None:
        return None
    return v if len(v) else None


def _upload(cloud, name):
    """What goes to the device for `name`: a cloud may carry the array in a narrower scalar type than the float64 view
    its getters give (io.FileCloud: raw_points / raw_colors / raw_normals) -- every widening on the device is exact."""
    v = getattr(cloud, "raw_" + name, None)
    if v is not None and len(v):
        return v
    return _attr(cloud, name)


class FusedDirection(typing.NamedTuple):
    n: int
    sum_d1: float          # exact for integer clouds
    max_d1: float
    sum_d2: float
    max_d2: float
    d2_valid: bool
    color_sum: np.ndarray
    color_max: np.ndarray


class CloudPair:
    def __init__(self, origin_cloud, reconst_cloud, *, ctx: N.Context | None = None,
                 normals_mode: str = "reference", peak: str = "obb", resolution_bits: int | None = None,
                 eager_normals: bool = False, knn: int = 30, cell_size: float = 0.0,
                 rank: int = 0, world: int = 1, group=None, ties: str = "first"):
        if normals_mode not in ("reference", "neighbour"):
            raise ValueError("normals_mode must be 'reference' or 'neighbour'")
        if peak not in ("obb", "aabb_diag", "resolution"):
            raise ValueError("peak must be 'obb', 'aabb_diag' or 'resolution'")
        if ties not in ("first", "average"):
            raise ValueError("ties must be 'first' or 'average'")
        self._ties = ties
        self.clouds = (origin_cloud, reconst_cloud)
        self._ctx = ctx or default_context()
        self._normals_mode = normals_mode
        self._peak = peak
        self._bits = resolution_bits
        self._knn = knn
        self._rank, self._world, self._group = rank, world, group
        self._fused_cache = {}
        self._idx = [None, None]
        self._d2 = [None, None]
        self._boundary = None
        self._extent = None
        self._normals_host = {}

        # upload + index (replaces the two KDTreeFlann builds, cloud_pair.py:65)
        # coordinates of BOTH clouds first: statistics and the index build need nothing else, colours and
        # normals follow on the copy stream while the index is built (they are first read by the epilogue)
        self._dev = []
        if world > 1:
            # one pair over several GPUs: integer pairs are split by slabs of z inside the library (this rank indexes and
            # queries its slab only); the partial sums are exchanged below with torch.distributed
            self._ctx.set_shard(rank, world)
        try:
            for c in self.clouds:
                self._dev.append(self._ctx.cloud(_upload(c, "points") if _upload(c, "points") is not None else np.zeros((0, 3))))
            for c, d in zip(self.clouds, self._dev):
                d.attach(_upload(c, "colors"), _upload(c, "normals"))
            self._ctx.build_pair(self._dev[0], self._dev[1], cell_size)
        except Exception:
            self.close()
            raise
        finally:
            if world > 1:
                self._ctx.set_shard(0, 1)
        # everything above is only ENQUEUED (uploads, statistics, index build): what the host has to wait for -- coordinate
        # kind, bounding boxes -- is fetched on first use, so that a caller can prepare the next pair in the meantime
        self._infos = None
        self._n = tuple(0 if _upload(c, "points") is None else len(_upload(c, "points")) for c in self.clouds)
        self._has_normals = [_upload(c, "normals") is not None for c in self.clouds]
        self._has_colors = tuple(_upload(c, "colors") is not None for c in self.clouds)
        if eager_normals:
            for k in range(2):
                self._ensure_normals(k)
        # the reference runs both NN passes in its constructor (cloud_pair.py:67-78) and
        # fails there on an empty search cloud; keep that behaviour
        if min(self._n) == 0:
            self.close()
            raise IndexError("list index out of range")

    def _info(self):
        if self._infos is None:
            self._infos = [d.info() for d in self._dev]
        return self._infos

    @property
    def kind(self):
        return self._info()[0].index_kind

    @property
    def _sharded(self):
        return bool(self._info()[0].sharded)

    @property
    def _aabb(self):
        return tuple((np.array(i.aabb_min), np.array(i.aabb_max)) for i in self._info())

    # ---- reference surface ---------------------------------------------------------
    @property
    def origin_cloud(self):
        return self.clouds[0]

    @property
    def reconst_cloud(self):
        return self.clouds[1]

    def get_left_error_vector(self):
        return self._error_vector(0)

    def get_right_error_vector(self):
        return self._error_vector(1)

    def get_left_neighbour_distances(self):
        self._materialise()
        return self._d2[0]

    def get_right_neighbour_distances(self):
        self._materialise()
        return self._d2[1]

    def get_boundary_sqrt_distances(self):
        n = self._n[0]
        if n < 2:
            return np.zeros((n,), dtype=np.float64)
        return self._dev[0].self_nn_minmax(per_point=True)[2]

    def get_extent(self):
        if self._extent is None:
            lo, hi = self._aabb[0]
            if self._peak == "obb":
                self._extent = _obb.minimal_obb_extent(np.asarray(self.clouds[0].points, dtype=np.float64), ctx=self._ctx,
                                                       dev_cloud=self._dev[0])
            elif self._peak == "aabb_diag":
                self._extent = np.full(3, _obb.aabb_diag(lo, hi))
            else:
                self._extent = np.full(3, _obb.resolution_peak(hi, self._bits))
        return self._extent

    def get_left_colors(self):
        return self.clouds[0].colors

    def get_right_colors(self):
        return self.clouds[1].colors

    def get_left_neighbour_colors(self):
        return self._neighbour_colors(0)

    def get_right_neighbour_colors(self):
        return self._neighbour_colors(1)

    # ---- additions --------------------------------------------------------------------
    def get_normals(self, k: int) -> np.ndarray:
        """Normals of cloud k (estimated on the GPU when the cloud came without)."""
        self._ensure_normals(k)
        nrm = _attr(self.clouds[k], "normals")
        if nrm is None:
            nrm = self._normals_host[k]
        return np.asarray(nrm)

    def get_neighbour_indices(self, is_left: bool) -> np.ndarray:
        self._materialise()
        return self._idx[0 if is_left else 1]

    def boundary_minmax(self):
        """(min, max) of get_boundary_sqrt_distances() reduced on the GPU."""
        if self._boundary is None:
            n = self._n[0]
            if n == 0:
                raise ValueError("zero-size array to reduction operation minimum which has no identity")
            if n < 2:
                self._boundary = (np.float64(0.0), np.float64(0.0))
            else:
                n0 = self._n[0]
                if self._sharded:      # split pair: the library reduces this rank's slab whatever range it is given
                    b, e = 0, n0
                else:
                    b, e = n0 * self._rank // self._world, n0 * (self._rank + 1) // self._world
                mn, mx, _ = self._dev[0].self_nn_minmax(b, e)
                if self._world > 1:
                    mn, mx = self._exchange_minmax(mn, mx)
                self._boundary = (np.float64(mn), np.float64(mx))
        return self._boundary

    def fused(self, is_left: bool, point_to_plane: bool = False, color_scheme: str | None = None) -> FusedDirection:
        """Reductions of one direction computed on the GPU: everything GeoMSE,
        GeoHausdorffDistance, ColorMSE and ColorHausdorffDistance need."""
        key = (bool(point_to_plane), color_scheme)
        # a richer evaluation already cached also answers a poorer request
        for (p2p, cs), res in self._fused_cache.items():
            if (p2p or not point_to_plane) and (cs == color_scheme or color_scheme is None):
                return res[0 if is_left else 1]
        flags = 0
        T = None
        scale = 1.0
        if point_to_plane:
            flags |= N.EVAL_D2
            self._ensure_normals(0)
            self._ensure_normals(1)
        if color_scheme is not None:
            if color_scheme not in _COLOR_TRANSFORMS:
                raise KeyError(color_scheme)
            if not all(self._has_colors):
                raise ValueError("colour metrics need colours on both clouds")
            flags |= N.EVAL_COLOR
            T = np.array(_COLOR_TRANSFORMS[color_scheme], dtype=np.float64)
            scale = 255.0 if color_scheme == "rgb" else 1.0  # metric.py:421-424
        if self._ties == "average" and flags:
            flags |= N.EVAL_TIE_AVERAGE
        mode = N.NORMALS_BY_QUERY_INDEX if self._normals_mode == "reference" else N.NORMALS_BY_NEIGHBOUR
        raw = self._ctx.pair_eval(self._dev[0], self._dev[1], flags, T, scale, mode, self._rank, self._world)
        dirs = [raw.dir[0], raw.dir[1]]
        vals = []
        for d in dirs:
            vals.append(dict(n=int(d.n_total), exact=bool(d.d1_exact_int), sum_u64=int(d.sum_d1_u64),
                             sum_d1=float(d.sum_d1), max_d1=float(d.max_d1), sum_d2=float(d.sum_d2),
                             max_d2=float(d.max_d2), d2_valid=bool(d.d2_valid),
                             csum=np.array(list(d.color_sum)), cmax=np.array(list(d.color_max))))
        if self._world > 1:
            vals = self._exchange_partials(vals)
        out = []
        for v in vals:
            s1 = float(v["sum_u64"]) if v["exact"] else v["sum_d1"]
            out.append(FusedDirection(v["n"], s1, v["max_d1"], v["sum_d2"], v["max_d2"], v["d2_valid"],
                                      v["csum"], v["cmax"]))
        self._fused_cache[key] = tuple(out)
        return out[0 if is_left else 1]

    # ---- internals ----------------------------------------------------------------------
    def _ensure_normals(self, k: int):
        if self._has_normals[k]:
            if _attr(self.clouds[k], "normals") is None:  # estimated before, write back once
                self._write_back_normals(k)
            return
        dev = self._dev[k]
        n = self._n[k]
        if self._world > 1 and n:
            import torch
            from . import distributed as D
            b, e = D.slice_range(n, self._rank, self._world)
            dev.estimate_normals(self._knn, b, e)  # buffer is zero outside [b, e)
            t = torch.empty((n, 3), dtype=torch.float64, device=f"cuda:{self._ctx.device}")
            dev.get_normals(t)
            self._ctx.synchronize()
            D.combine_disjoint(t, self._group)
            torch.cuda.synchronize()
            dev.set_normals(t)
            self._ctx.synchronize()
        else:
            dev.estimate_normals(self._knn)
        self._has_normals[k] = True
        self._write_back_normals(k)

    def _write_back_normals(self, k: int):
        nrm = self._dev[k].get_normals()
        try:
            self.clouds[k].normals = nrm  # cloud_pair.py:61-64 fills the caller's object in place
        except Exception:  # read-only foreign object: keep a private copy
            pass
        self._normals_host[k] = nrm

    def _materialise(self):
        if self._idx[0] is not None:
            return
        if self._world > 1:
            raise RuntimeError("per-point getters are single-process only; use fused() under world > 1")
        self._ctx.pair_eval(self._dev[0], self._dev[1], N.EVAL_PERPOINT)
        for d in range(2):
            self._idx[d] = self._ctx.pair_get(N.GET_IDX, d, self._n[d]).astype(np.int64)  # cloud_pair.py:33
            self._d2[d] = self._ctx.pair_get(N.GET_D2, d, self._n[d])

    def _error_vector(self, d: int):
        self._materialise()
        q = np.asarray(self.clouds[d].points, dtype=np.float64)
        s = np.asarray(self.clouds[1 - d].points, dtype=np.float64)
        return np.subtract(q, np.take(s, self._idx[d], axis=0))   # cloud_pair.py:34,90-100

    def _neighbour_colors(self, d: int):
        self._materialise()
        cols = _attr(self.clouds[1 - d], "colors")
        if cols is None:
            return np.zeros((0, 3))
        return np.take(np.asarray(cols, dtype=np.float64), self._idx[d], axis=0)   # cloud_pair.py:38-40

    # ---- multi-GPU exchange of the tiny partial records ------------
    def _exchange_partials(self, vals):
        from . import distributed as D
        return D.exchange_partials(vals, self._world, self._group, f"cuda:{self._ctx.device}")

    def _exchange_minmax(self, mn, mx):
        from . import distributed as D
        return D.exchange_minmax(mn, mx, self._group, f"cuda:{self._ctx.device}")

    def close(self):
        for d in self._dev:
            d.close()
        self._dev = []
