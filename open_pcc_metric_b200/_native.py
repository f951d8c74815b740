"""ctypes binding of libpccm.so (C ABI: include/pccm.h).

There is deliberately no fallback: if the CUDA library is missing or no GPU is
present every entry point raises.  The library is built in-tree by
``__graft_entry__.build()`` / ``make -C open_pcc_metric_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PCCM_LIB", os.path.join(_HERE, "libpccm.so"))  # PCCM_LIB: A/B testing of builds

# enums (include/pccm.h)
F64, F32, I32, U16, U8 = 0, 1, 2, 3, 4
HOST, DEVICE = 0, 1
KIND_AUTO, KIND_INT, KIND_F32, KIND_F64 = -1, 0, 1, 2
EVAL_D2, EVAL_COLOR, EVAL_PERPOINT, EVAL_TIE_AVERAGE = 1, 2, 4, 8
NORMALS_BY_QUERY_INDEX, NORMALS_BY_NEIGHBOUR = 0, 1
GET_IDX, GET_D2 = 0, 1
ERR_INVALID, ERR_CUDA, ERR_STATE, ERR_INDEX, ERR_NONFINITE, ERR_UNSUPPORTED = -1, -2, -3, -4, -5, -6

EXPORTS = [
    "pccm_version", "pccm_last_error", "pccm_ctx_create", "pccm_ctx_destroy", "pccm_ctx_synchronize",
    "pccm_ctx_set_profiling", "pccm_ctx_reset_timings", "pccm_ctx_get_timings", "pccm_ctx_set_shard",
    "pccm_cloud_create", "pccm_cloud_attach", "pccm_cloud_destroy", "pccm_cloud_info_get", "pccm_cloud_build_index", "pccm_pair_build_index",
    "pccm_cloud_set_normals", "pccm_cloud_get_normals", "pccm_estimate_normals", "pccm_knn_self",
    "pccm_self_nn_minmax", "pccm_nn", "pccm_pair_eval", "pccm_pair_get", "pccm_obb_sweep", "pccm_cloud_extremes", "pccm_cloud_outside_hull",
]


class CloudInfo(C.Structure):
    _fields_ = [("n", C.c_int64), ("data_kind", C.c_int32), ("index_kind", C.c_int32),
                ("has_colors", C.c_int32), ("colors_u8", C.c_int32), ("has_normals", C.c_int32),
                ("indexed", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32), ("sharded", C.c_int32), ("reserved", C.c_int32),
                ("cell_size", C.c_double),
                ("aabb_min", C.c_double * 3), ("aabb_max", C.c_double * 3)]


class DirResult(C.Structure):
    _fields_ = [("n", C.c_int64), ("n_total", C.c_int64), ("sum_d1_u64", C.c_uint64),
                ("d1_exact_int", C.c_int32), ("d2_valid", C.c_int32),
                ("sum_d1", C.c_double), ("max_d1", C.c_double), ("sum_d2", C.c_double), ("max_d2", C.c_double),
                ("color_sum", C.c_double * 3), ("color_max", C.c_double * 3)]


class PairResult(C.Structure):
    _fields_ = [("dir", DirResult * 2)]


class Timings(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("upload_ms", "stats_ms", "keys_ms", "sort_ms", "table_ms",
                                           "reorder_ms", "query_ms", "finalize_ms", "knn_ms",
                                           "vox_build_ms", "vox_tail_ms", "vox_search_ms", "vox_epilogue_ms")] + \
               [(k, C.c_int64) for k in ("query_launches", "knn_launches", "total_launches", "library_launches",
                                         "vox_undecided", "vox_far", "vox_tail")]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class PccmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libpccm error {code}: {msg}")
        self.code = code
        self.msg = msg


_lib = None


def lib():
    """Load libpccm.so; raise loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()' "
            "or make -C open_pcc_metric_b200/csrc).  open_pcc_metric_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    L.pccm_last_error.restype = C.c_char_p
    L.pccm_last_error.argtypes = [C.c_void_p]
    L.pccm_version.restype = C.c_int
    vp, i32, i64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_double
    sig = {
        "pccm_ctx_create": [i32, vp, C.POINTER(vp)],
        "pccm_ctx_destroy": [vp],
        "pccm_ctx_synchronize": [vp],
        "pccm_ctx_set_profiling": [vp, i32],
        "pccm_ctx_reset_timings": [vp],
        "pccm_ctx_set_shard": [vp, i32, i32],
        "pccm_ctx_get_timings": [vp, C.POINTER(Timings)],
        "pccm_cloud_create": [vp, vp, i32, i64, i64, vp, i32, i64, vp, i32, i64, i32, C.POINTER(vp)],
        "pccm_cloud_attach": [vp, vp, vp, i32, i64, vp, i32, i64, i32],
        "pccm_cloud_destroy": [vp, vp],
        "pccm_cloud_info_get": [vp, vp, C.POINTER(CloudInfo)],
        "pccm_cloud_build_index": [vp, vp, dbl, i32],
        "pccm_pair_build_index": [vp, vp, vp, dbl, i32],
        "pccm_cloud_set_normals": [vp, vp, vp, i32, i64, i32],
        "pccm_cloud_get_normals": [vp, vp, vp, i32],
        "pccm_estimate_normals": [vp, vp, i32, i64, i64],
        "pccm_knn_self": [vp, vp, i32, vp, vp, i32],
        "pccm_self_nn_minmax": [vp, vp, i64, i64, C.POINTER(dbl), C.POINTER(dbl), vp, i32],
        "pccm_nn": [vp, vp, vp, vp, vp, i32],
        "pccm_pair_eval": [vp, vp, vp, C.c_uint32, vp, dbl, i32, i32, i32, C.POINTER(PairResult)],
        "pccm_pair_get": [vp, i32, i32, vp, i32],
        "pccm_obb_sweep": [vp, vp, i64, vp, i64, vp, vp],
        "pccm_cloud_extremes": [vp, vp, vp, i32, vp],
        "pccm_cloud_outside_hull": [vp, vp, vp, i32, dbl, i64, vp, C.POINTER(i64)],
    }
    for name, args in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = C.c_int
    _lib = L
    return L


_NP_DTYPES = {np.dtype(np.float64): F64, np.dtype(np.float32): F32, np.dtype(np.int32): I32,
              np.dtype(np.uint16): U16, np.dtype(np.uint8): U8}


def _is_torch(x):
    return type(x).__module__.startswith("torch") and hasattr(x, "data_ptr")


_TORCH_MAP = None


def _torch_map():
    global _TORCH_MAP
    if _TORCH_MAP is None:
        import torch
        _TORCH_MAP = {torch.float64: (F64, 8), torch.float32: (F32, 4), torch.int32: (I32, 4), torch.uint8: (U8, 1)}
        if hasattr(torch, "uint16"):
            _TORCH_MAP[torch.uint16] = (U16, 2)
    return _TORCH_MAP


class _Buf:
    """(pointer, dtype code, row stride in bytes, mem kind) + a reference keeping it alive."""

    def __init__(self, arr, allowed, what):
        if _is_torch(arr):
            ent = _torch_map().get(arr.dtype)
            if ent is None:
                raise TypeError(f"{what}: unsupported torch dtype {arr.dtype}")
            shape, st = arr.shape, arr.stride()
            if len(shape) != 2 or shape[1] < 3 or st[1] != 1:
                raise ValueError(f"{what}: expected (N, 3) rows with unit inner stride")
            self.keep = arr
            self.ptr = arr.data_ptr()
            self.dtype, es = ent
            self.stride = st[0] * es if shape[0] > 1 else 3 * es
            self.mem = DEVICE if arr.is_cuda else HOST
            self.n = shape[0]
        else:
            a = np.asarray(arr)
            if a.dtype not in _NP_DTYPES:
                a = a.astype(np.float64)
            if a.ndim != 2 or a.shape[1] != 3:
                raise ValueError(f"{what}: expected shape (N, 3), got {a.shape}")
            if a.shape[0] and (a.strides[1] != a.itemsize or a.strides[0] < 3 * a.itemsize):
                a = np.ascontiguousarray(a)
            self.keep = a
            self.ptr = a.ctypes.data
            self.dtype = _NP_DTYPES[a.dtype]
            self.stride = a.strides[0] if a.shape[0] > 1 else 3 * a.itemsize
            self.mem = HOST
            self.n = a.shape[0]
        if self.dtype not in allowed:
            raise TypeError(f"{what}: dtype code {self.dtype} not accepted here")


class Context:
    """One pccm_ctx (one CUDA stream on one device)."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self.L = lib()
        h = C.c_void_p()
        rc = self.L.pccm_ctx_create(int(device), C.c_void_p(stream or 0), C.byref(h))
        if rc != 0:
            raise PccmError(rc, (self.L.pccm_last_error(None) or b"").decode())
        self.h = h
        self.device = device

    def check(self, rc):
        if rc == 0:
            return
        msg = (self.L.pccm_last_error(self.h) or b"").decode()
        if rc == ERR_INDEX:
            raise IndexError(msg)
        if rc == ERR_INVALID:
            raise ValueError(msg)
        raise PccmError(rc, msg)

    def close(self):
        if getattr(self, "h", None):
            self.L.pccm_ctx_destroy(self.h)
            self.h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        self.check(self.L.pccm_ctx_synchronize(self.h))

    def set_profiling(self, level: int):
        """0 = off, 1 = query / k-NN kernels only, 2 = every stage."""
        self.check(self.L.pccm_ctx_set_profiling(self.h, int(level)))

    def set_shard(self, rank: int = 0, world: int = 1):
        """Integer pairs built from now on are split over ``world`` ranks by slabs of z (see pccm_ctx_set_shard)."""
        self.check(self.L.pccm_ctx_set_shard(self.h, int(rank), int(world)))

    def reset_timings(self):
        self.check(self.L.pccm_ctx_reset_timings(self.h))

    def timings(self) -> dict:
        t = Timings()
        self.check(self.L.pccm_ctx_get_timings(self.h, C.byref(t)))
        return t.as_dict()

    def cloud(self, points, colors=None, normals=None) -> "Cloud":
        return Cloud(self, points, colors, normals)

    # --- pair level ---------------------------------------------------------
    def build_pair(self, a: "Cloud", b: "Cloud", cell_size: float = 0.0, force_kind: int = KIND_AUTO):
        """Index both clouds of a pair in joint launches (common coordinate kind)."""
        self.check(self.L.pccm_pair_build_index(self.h, a.h, b.h, float(cell_size), int(force_kind)))
        a._keep = b._keep = None

    def nn(self, query: "Cloud", search: "Cloud"):
        n = query.n
        idx = np.empty(n, dtype=np.int32)
        d2 = np.empty(n, dtype=np.float64)
        self.check(self.L.pccm_nn(self.h, query.h, search.h, idx.ctypes.data, d2.ctypes.data, HOST))
        return idx, d2

    def pair_eval(self, a: "Cloud", b: "Cloud", flags=0, color_matrix=None, color_scale=1.0,
                  normals_mode=NORMALS_BY_QUERY_INDEX, rank=0, world=1) -> PairResult:
        res = PairResult()
        T = None
        if color_matrix is not None:
            T = np.ascontiguousarray(color_matrix, dtype=np.float64).reshape(9)
        self.check(self.L.pccm_pair_eval(self.h, a.h, b.h, int(flags), T.ctypes.data if T is not None else None,
                                         float(color_scale), int(normals_mode), int(rank), int(world), C.byref(res)))
        return res

    def obb_sweep(self, hull_vertices, triangles):
        """(volume[nf], extent[nf, 3]) of the hull vertices' box in every hull triangle's frame."""
        hv = np.ascontiguousarray(hull_vertices, dtype=np.float64)
        tr = np.ascontiguousarray(triangles, dtype=np.float64).reshape(-1, 9)
        vol = np.empty(len(tr), dtype=np.float64)
        ext = np.empty((len(tr), 3), dtype=np.float64)
        self.check(self.L.pccm_obb_sweep(self.h, hv.ctypes.data, len(hv), tr.ctypes.data, len(tr), vol.ctypes.data, ext.ctypes.data))
        return vol, ext

    def pair_get(self, which, direction, n):
        out = np.empty(n, dtype=np.int32 if which == GET_IDX else np.float64)
        self.check(self.L.pccm_pair_get(self.h, which, direction, out.ctypes.data, HOST))
        return out


class Cloud:
    """One pccm_cloud handle."""

    def __init__(self, ctx: Context, points, colors=None, normals=None):
        self.ctx = ctx
        self.h = None
        pb = _Buf(points, (F64, F32, I32, U16), "points")
        cb = _Buf(colors, (F64, U8), "colors") if colors is not None and len(colors) else None
        nb = _Buf(normals, (F64, F32), "normals") if normals is not None and len(normals) else None
        for b, what in ((cb, "colors"), (nb, "normals")):
            if b is not None and (b.n != pb.n or b.mem != pb.mem):
                raise ValueError(f"{what}: must match points in length and memory kind")
        h = C.c_void_p()
        ctx.check(ctx.L.pccm_cloud_create(
            ctx.h, pb.ptr, pb.dtype, pb.n, pb.stride,
            cb.ptr if cb else None, cb.dtype if cb else F64, cb.stride if cb else 0,
            nb.ptr if nb else None, nb.dtype if nb else F64, nb.stride if nb else 0,
            pb.mem, C.byref(h)))
        self.h = h
        self.n = pb.n
        self._keep = pb            # coordinates are read until the index is built
        self._keep_attr = (cb, nb)  # colours / normals travel on the copy stream (pinned host inputs) or are used in place (device normals)
        self._keep_normals = nb   # device normals are used in place; pinned host normals may be in flight on the copy stream

    def attach(self, colors=None, normals=None):
        """Colours and / or normals for a cloud created from coordinates only (uploaded on the copy
        stream beside the index build; see pccm_cloud_attach)."""
        cb = _Buf(colors, (F64, U8), "colors") if colors is not None and len(colors) else None
        nb = _Buf(normals, (F64, F32), "normals") if normals is not None and len(normals) else None
        if cb is None and nb is None:
            return
        mem = (cb or nb).mem
        for b, what in ((cb, "colors"), (nb, "normals")):
            if b is not None and (b.n != self.n or b.mem != mem):
                raise ValueError(f"{what}: must match the cloud in length and share one memory kind")
        self.ctx.check(self.ctx.L.pccm_cloud_attach(
            self.ctx.h, self.h, cb.ptr if cb else None, cb.dtype if cb else F64, cb.stride if cb else 0,
            nb.ptr if nb else None, nb.dtype if nb else F64, nb.stride if nb else 0, mem))
        self._keep_attr = (cb, nb)     # pinned host inputs may be in flight, device normals are used in place

    def info(self) -> CloudInfo:
        out = CloudInfo()
        self.ctx.check(self.ctx.L.pccm_cloud_info_get(self.ctx.h, self.h, C.byref(out)))
        return out

    def build_index(self, cell_size: float = 0.0, force_kind: int = KIND_AUTO):
        self.ctx.check(self.ctx.L.pccm_cloud_build_index(self.ctx.h, self.h, float(cell_size), int(force_kind)))
        self._keep = None

    def set_normals(self, normals):
        nb = _Buf(normals, (F64, F32), "normals")
        if nb.n != self.n:
            raise ValueError("normals: wrong length")
        self.ctx.check(self.ctx.L.pccm_cloud_set_normals(self.ctx.h, self.h, nb.ptr, nb.dtype, nb.stride, nb.mem))

    def get_normals(self, out=None):
        if out is not None and _is_torch(out):
            self.ctx.check(self.ctx.L.pccm_cloud_get_normals(self.ctx.h, self.h, out.data_ptr(), DEVICE))
            return out
        out = np.empty((self.n, 3), dtype=np.float64)
        self.ctx.check(self.ctx.L.pccm_cloud_get_normals(self.ctx.h, self.h, out.ctypes.data, HOST))
        return out

    def estimate_normals(self, k: int = 30, begin: int = 0, end: int | None = None):
        self.ctx.check(self.ctx.L.pccm_estimate_normals(self.ctx.h, self.h, int(k), int(begin),
                                                        int(self.n if end is None else end)))

    def knn_self(self, k: int):
        idx = np.empty((self.n, k), dtype=np.int32)
        d2 = np.empty((self.n, k), dtype=np.float64)
        self.ctx.check(self.ctx.L.pccm_knn_self(self.ctx.h, self.h, int(k), idx.ctypes.data, d2.ctypes.data, HOST))
        return idx, d2

    def self_nn_minmax(self, begin: int = 0, end: int | None = None, per_point: bool = False):
        mn, mx = C.c_double(), C.c_double()
        pp = np.empty(self.n, dtype=np.float64) if per_point else None
        self.ctx.check(self.ctx.L.pccm_self_nn_minmax(self.ctx.h, self.h, int(begin), int(self.n if end is None else end),
                                                      C.byref(mn), C.byref(mx), pp.ctypes.data if per_point else None, HOST))
        return mn.value, mx.value, pp

    def extremes(self, dirs):
        """Original indices of the arg-max of dirs[d] . p."""
        d = np.ascontiguousarray(dirs, dtype=np.float64).reshape(-1, 3)
        out = np.empty(len(d), dtype=np.int32)
        self.ctx.check(self.ctx.L.pccm_cloud_extremes(self.ctx.h, self.h, d.ctypes.data, len(d), out.ctypes.data))
        return out

    def outside_hull(self, planes, eps: float, capacity: int | None = None):
        """Points not strictly inside the polytope planes[f] = (n, d); returns an (M, 3) array."""
        pl = np.ascontiguousarray(planes, dtype=np.float64).reshape(-1, 4)
        cap = self.n if capacity is None else int(capacity)
        out = np.empty((cap, 3), dtype=np.float64)
        cnt = C.c_int64()
        self.ctx.check(self.ctx.L.pccm_cloud_outside_hull(self.ctx.h, self.h, pl.ctypes.data, len(pl), float(eps), cap,
                                                          out.ctypes.data, C.byref(cnt)))
        return out[:min(cnt.value, cap)], cnt.value

    def close(self):
        if self.h is not None and self.ctx.h is not None:
            self.ctx.L.pccm_cloud_destroy(self.ctx.h, self.h)
        self.h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass
