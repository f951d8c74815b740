"""Point-cloud file reader standing in for ``o3d.io.read_point_cloud``
(reference handler.py:57).

PLY (ascii, binary_little_endian, binary_big_endian) with x/y/z of any scalar type,
optional red/green/blue (uchar -> /255, floats kept) and nx/ny/nz; plus whitespace
separated ``.xyz`` / ``.txt`` (x y z [r g b]).  Returns a ``geometry.PointCloud``
holding float64 arrays like Open3D does -- or, with ``native=True``, a ``FileCloud`` that
keeps the file's own scalar types for the upload (a voxelised PLY with float / ushort
coordinates, uchar colours and float normals is 21 bytes per point instead of the 72 of
three float64 rows: the evaluation from host arrays is bound by PCIe, bench.py
``e2e.compact_inputs``) and converts to float64 only when a getter asks.  Every widening the
library then does on the device (uint16 / int32 / float32 -> float64, uchar -> k / 255.0) is
exact, so the metrics are those of the float64 copy.
"""
from __future__ import annotations

import numpy as np

from .geometry import PointCloud

_PLY_TYPES = {
    "char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2",
    "ushort": "u2", "uint16": "u2", "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4",
    "float": "f4", "float32": "f4", "double": "f8", "float64": "f8",
}


_UPLOAD_DTYPES = ("f8", "f4", "i4", "u2")      # what pccm_cloud_create takes for coordinates (include/pccm.h; uchar coordinates are widened here)


class FileCloud:
    """A cloud as its file holds it.  ``raw_points`` / ``raw_colors`` / ``raw_normals`` are what ``CloudPair`` uploads
    ((N, 3) arrays in the file's scalar types; uchar colours stay 0..255); ``points`` / ``colors`` / ``normals`` are the
    float64 arrays Open3D would hold (colours in [0, 1]), made on first use."""

    def __init__(self, raw_points, raw_colors=None, raw_normals=None):
        self.raw_points = raw_points
        self.raw_colors = raw_colors
        self.raw_normals = raw_normals
        self._f64 = {}

    def _view(self, name, raw, scale=None):
        if raw is None:
            return None
        if name not in self._f64:
            a = np.asarray(raw, dtype=np.float64)
            self._f64[name] = a / scale if scale else a
        return self._f64[name]

    @property
    def points(self):
        return self._view("points", self.raw_points)

    @property
    def colors(self):
        c = self.raw_colors
        return self._view("colors", c, 255.0 if c is not None and c.dtype == np.uint8 else None)

    @property
    def normals(self):
        return self._view("normals", self.raw_normals)

    @normals.setter
    def normals(self, v):                       # (CloudPair writes estimated normals back into the caller's cloud)
        self.raw_normals = None if v is None else np.asarray(v, dtype=np.float64)
        self._f64.pop("normals", None)

    def __len__(self):
        return len(self.raw_points)

    def has_colors(self):
        return self.raw_colors is not None and len(self.raw_colors) == len(self.raw_points) > 0

    def has_normals(self):
        return self.raw_normals is not None and len(self.raw_normals) == len(self.raw_points) > 0


def _pinned_like(shape, dtype):
    """A page-locked (N, 3) host array when torch + CUDA are there (uploads from it are truly asynchronous: CloudPair only
    enqueues), else an ordinary one."""
    try:
        import torch
        tdt = {"f8": torch.float64, "f4": torch.float32, "i4": torch.int32, "u2": getattr(torch, "uint16", None), "u1": torch.uint8}[np.dtype(dtype).str[1:]]
        if tdt is not None and torch.cuda.is_available():
            return torch.empty(shape, dtype=tdt, pin_memory=True).numpy()
    except Exception:
        pass
    return np.empty(shape, dtype=dtype)


def _read_ply(path: str, native: bool = False, pinned: bool = False):
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt = None
        elements = []  # (name, count, [(prop, type) | (prop, 'list', count_t, item_t)])
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: truncated header")
            tok = line.decode("ascii", "replace").split()
            if not tok or tok[0] == "comment" or tok[0] == "obj_info":
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                elements.append((tok[1], int(tok[2]), []))
            elif tok[0] == "property":
                if tok[1] == "list":
                    elements[-1][2].append((tok[4], "list", tok[2], tok[3]))
                else:
                    elements[-1][2].append((tok[2], tok[1]))
            elif tok[0] == "end_header":
                break
        if fmt not in ("ascii", "binary_little_endian", "binary_big_endian"):
            raise ValueError(f"{path}: unsupported PLY format {fmt}")
        vertex = None
        for name, count, props in elements:
            if name != "vertex":
                if vertex is None:
                    raise ValueError(f"{path}: element '{name}' precedes 'vertex' (unsupported)")
                break
            if any(p[1] == "list" for p in props):
                raise ValueError(f"{path}: list property inside vertex element")
            names = [p[0] for p in props]
            ftype = {p[0]: _PLY_TYPES[p[1]] for p in props}           # scalar type of every property, as declared
            if fmt == "ascii":
                data = np.loadtxt(f, dtype=np.float64, max_rows=count, ndmin=2) if count else np.zeros((0, len(names)))
                vertex = {n: data[:, i] for i, n in enumerate(names)}
                u8 = {p[0] for p in props if _PLY_TYPES[p[1]] == "u1"}
            else:
                end = "<" if fmt == "binary_little_endian" else ">"
                dt = np.dtype([(p[0], end + _PLY_TYPES[p[1]]) for p in props])
                raw = np.frombuffer(f.read(dt.itemsize * count), dtype=dt, count=count)
                vertex = {n: raw[n] for n in names}
                u8 = {p[0] for p in props if _PLY_TYPES[p[1]] == "u1"}
        if vertex is None:
            raise ValueError(f"{path}: no vertex element")

    def cols(keys):
        return np.stack([np.asarray(vertex[k], dtype=np.float64) for k in keys], axis=1)

    if native:
        def raw(keys, allowed):
            # one array in the file's scalar type when the three properties share one the library uploads; text files
            # only keep INTEGER types (a decimal fraction is the float64 the default reader parses, not its float32)
            t = {ftype[k] for k in keys}
            dt = t.pop() if len(t) == 1 else None
            if not (dt in allowed and (fmt != "ascii" or dt[0] in "iu")):
                dt = "f8"
            out = _pinned_like((len(vertex[keys[0]]), 3), dt) if pinned else np.empty((len(vertex[keys[0]]), 3), dtype=dt)
            for j, k in enumerate(keys):                  # (interleaves the file's columns straight into the upload array)
                out[:, j] = vertex[k]
            return out
        fc = FileCloud(raw(("x", "y", "z"), _UPLOAD_DTYPES))
        if all(k in vertex for k in ("red", "green", "blue")):
            fc.raw_colors = raw(("red", "green", "blue"), ("u1", "f8"))
        if all(k in vertex for k in ("nx", "ny", "nz")):
            fc.raw_normals = raw(("nx", "ny", "nz"), ("f4", "f8"))
        return fc
    pc = PointCloud(cols(("x", "y", "z")))
    if all(k in vertex for k in ("red", "green", "blue")):
        c = cols(("red", "green", "blue"))
        if "red" in u8:
            c = c / 255.0
        pc.colors = c
    if all(k in vertex for k in ("nx", "ny", "nz")):
        pc.normals = cols(("nx", "ny", "nz"))
    return pc


def _read_xyz(path: str) -> PointCloud:
    data = np.loadtxt(path, dtype=np.float64, ndmin=2)
    pc = PointCloud(data[:, :3])
    if data.shape[1] >= 6:
        c = data[:, 3:6]
        pc.colors = c / 255.0 if c.max() > 1.0 else c
    return pc


def read_point_cloud(path: str, native: bool = False, pinned: bool = False):
    """native=True (PLY): a FileCloud in the file's scalar types instead of a float64 PointCloud (see the module text);
    pinned=True: its arrays are page-locked when a CUDA device is there."""
    low = path.lower()
    if low.endswith(".ply"):
        return _read_ply(path, native, pinned)
    if low.endswith((".xyz", ".txt", ".xyzrgb")):
        return _read_xyz(path)
    raise ValueError(f"unsupported point cloud format: {path}")


def write_ply(path: str, cloud, binary: bool = True) -> None:
    """Small writer used by tests and examples (uchar colours, float normals)."""
    pts = np.asarray(cloud.points, dtype=np.float64)
    n = len(pts)
    fields = [("x", "<f8"), ("y", "<f8"), ("z", "<f8")]
    has_c = getattr(cloud, "colors", None) is not None and len(cloud.colors) == n and n > 0
    has_n = getattr(cloud, "normals", None) is not None and len(cloud.normals) == n and n > 0
    if has_c:
        fields += [("red", "u1"), ("green", "u1"), ("blue", "u1")]
    if has_n:
        fields += [("nx", "<f8"), ("ny", "<f8"), ("nz", "<f8")]
    rec = np.zeros(n, dtype=np.dtype(fields))
    rec["x"], rec["y"], rec["z"] = pts[:, 0], pts[:, 1], pts[:, 2]
    if has_c:
        c = np.rint(np.asarray(cloud.colors) * 255.0).astype(np.uint8)
        rec["red"], rec["green"], rec["blue"] = c[:, 0], c[:, 1], c[:, 2]
    if has_n:
        nr = np.asarray(cloud.normals)
        rec["nx"], rec["ny"], rec["nz"] = nr[:, 0], nr[:, 1], nr[:, 2]
    names = {"<f8": "double", "u1": "uchar"}
    header = ["ply", "format binary_little_endian 1.0" if binary else "format ascii 1.0", f"element vertex {n}"]
    header += [f"property {names[t]} {k}" for k, t in fields]
    header.append("end_header")
    with open(path, "wb") as f:
        f.write(("\n".join(header) + "\n").encode("ascii"))
        if binary:
            f.write(rec.tobytes())
        else:
            for r in rec:
                f.write((" ".join(repr(v.item()) if isinstance(v, np.floating) else str(v) for v in r) + "\n").encode("ascii"))
