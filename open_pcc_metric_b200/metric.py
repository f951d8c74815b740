"""Metric classes -- same names, keys, dependency keywords and value types as the
reference's metric.py (/root/reference/open_pcc_metric/metric.py:14-485), so code
written against it keeps working.

Two ways to a value:

* the reference's dependency graph (``_get_dependencies`` + ``calculate(**deps)``),
  fed by CloudPair's getters -- per-point numpy arrays, used when somebody asks for
  a per-point quantity or passes hand-made dependencies (as the reference's own
  tests do, tests/unit/test_metric.py:34-70);
* ``calculate_fused(cloud_pair)`` on the reductions (GeoMSE, GeoHausdorffDistance,
  ColorMSE, ColorHausdorffDistance, BoundarySqrtDistances): the sums and maxima were
  already reduced on the GPU by the query kernel, nothing per-point reaches the host.

The per-row Python loops of the reference (metric.py:139-143, :148-152, :286-290)
are whole-array numpy expressions here; values agree to rounding (exactly, for
integer clouds).
"""
from __future__ import annotations

import abc
import typing

import numpy as np

from .cloud_pair import CloudPair, _COLOR_TRANSFORMS


class AbstractMetric(abc.ABC):
    value: typing.Any

    def _key(self) -> tuple:
        return (type(self).__name__,)

    @abc.abstractmethod
    def calculate(self, *args, **kwargs) -> None:
        raise NotImplementedError("calculate is not implemented")

    def __str__(self) -> str:
        return f"{self._key()}: {self.value}"


class PrimaryMetric(AbstractMetric):
    """Pulls its value straight from a CloudPair."""

    @abc.abstractmethod
    def calculate(self, cloud_pair: CloudPair) -> None:
        raise NotImplementedError("calculate is not implemented")


class SecondaryMetric(AbstractMetric):
    """Computed from other metrics (keyword names = keys of _get_dependencies)."""

    def _get_dependencies(self) -> typing.Dict[str, AbstractMetric]:
        return {}

    @abc.abstractmethod
    def calculate(self, **kwargs) -> None:
        raise NotImplementedError("calculate is not implemented")


class DirectionalMetric(AbstractMetric):
    def __init__(self, is_left: bool):
        self.is_left = is_left

    def _key(self) -> tuple:
        return super()._key() + (self.is_left,)


class PointToPlaneable(DirectionalMetric):
    def __init__(self, is_left: bool, point_to_plane: bool):
        super().__init__(is_left)
        self.point_to_plane = point_to_plane

    def _key(self) -> tuple:
        return super()._key() + (self.point_to_plane,)


class ColorMetric(DirectionalMetric):
    def __init__(self, is_left: bool, color_scheme: str):
        super().__init__(is_left)
        self.color_scheme = color_scheme

    def _key(self) -> tuple:
        return super()._key() + (self.color_scheme,)


def _side(pair, is_left, left_name, right_name):
    return getattr(pair, left_name if is_left else right_name)()


# ---- primary metrics (metric.py:74-121, 182-188) ---------------------------------------
class PrimaryErrorVector(PrimaryMetric, DirectionalMetric):
    def calculate(self, cloud_pair: CloudPair) -> None:
        self.value = _side(cloud_pair, self.is_left, "get_left_error_vector", "get_right_error_vector")


class NeighbourDistances(PrimaryMetric, DirectionalMetric):
    def calculate(self, cloud_pair: CloudPair) -> None:
        self.value = _side(cloud_pair, self.is_left, "get_left_neighbour_distances", "get_right_neighbour_distances")


class CloudNormals(PrimaryMetric, DirectionalMetric):
    def calculate(self, cloud_pair: CloudPair) -> None:
        k = 0 if self.is_left else 1
        if hasattr(cloud_pair, "get_normals"):
            self.value = cloud_pair.get_normals(k)       # estimates on the GPU when absent
        else:
            self.value = np.asarray(cloud_pair.clouds[k].normals)


class CloudExtent(PrimaryMetric):
    def calculate(self, cloud_pair: CloudPair) -> None:
        self.value = cloud_pair.get_extent()


class CloudColors(PrimaryMetric, DirectionalMetric):
    def calculate(self, cloud_pair: CloudPair) -> None:
        self.value = _side(cloud_pair, self.is_left, "get_left_colors", "get_right_colors")


class NeighbourColors(PrimaryMetric, DirectionalMetric):
    def calculate(self, cloud_pair: CloudPair) -> None:
        self.value = _side(cloud_pair, self.is_left, "get_left_neighbour_colors", "get_right_neighbour_colors")


class BoundarySqrtDistances(PrimaryMetric):
    def calculate(self, cloud_pair: CloudPair) -> None:
        if hasattr(cloud_pair, "boundary_minmax"):
            self.value = cloud_pair.boundary_minmax()    # min / max reduced on the GPU
            return
        inner = cloud_pair.get_boundary_sqrt_distances()
        self.value = (np.min(inner), np.max(inner))


# ---- geometry (metric.py:124-247, 353-386) ------------------------------------------------
class ErrorVector(SecondaryMetric, PointToPlaneable):
    def _get_dependencies(self):
        deps = {"primary_error_vector": PrimaryErrorVector(is_left=self.is_left)}
        if self.point_to_plane:
            # the OTHER cloud's normals, later indexed by the query index (quirk Q1)
            deps["cloud_normals"] = CloudNormals(is_left=not self.is_left)
        return deps

    def calculate(self, primary_error_vector, cloud_normals=None) -> None:
        e = np.asarray(primary_error_vector.value)
        if not self.point_to_plane:
            self.value = np.sqrt((e[:, 0] * e[:, 0] + e[:, 1] * e[:, 1]) + e[:, 2] * e[:, 2])
            return
        nrm = np.asarray(cloud_normals.value)
        rows = e.shape[0]
        if nrm.shape[0] < rows:
            raise IndexError(f"index {nrm.shape[0]} is out of bounds for axis 0 with size {nrm.shape[0]}")
        nrm = nrm[:rows]
        self.value = (e[:, 0] * nrm[:, 0] + e[:, 1] * nrm[:, 1]) + e[:, 2] * nrm[:, 2]


class EuclideanDistance(SecondaryMetric, PointToPlaneable):
    def _get_dependencies(self):
        if self.point_to_plane:
            return {"error_vector": ErrorVector(is_left=self.is_left, point_to_plane=True)}
        return {"neighbour_distances": NeighbourDistances(is_left=self.is_left)}

    def calculate(self, neighbour_distances=None, error_vector=None) -> None:
        if self.point_to_plane:
            self.value = np.square(error_vector.value)
        else:
            self.value = neighbour_distances.value      # squared already (cloud_pair.py:22-23)


class MinSqrtDistance(SecondaryMetric):
    def _get_dependencies(self):
        return {"boundary_metric": BoundarySqrtDistances()}

    def calculate(self, boundary_metric) -> None:
        self.value = boundary_metric.value[0]


class MaxSqrtDistance(SecondaryMetric):
    def _get_dependencies(self):
        return {"boundary_metric": BoundarySqrtDistances()}

    def calculate(self, boundary_metric) -> None:
        self.value = boundary_metric.value[1]


def _fused_dir(cloud_pair, is_left, point_to_plane=False, color_scheme=None):
    fd = cloud_pair.fused(is_left, point_to_plane=point_to_plane, color_scheme=color_scheme)
    if point_to_plane and not fd.d2_valid:
        # what metric.py:148-152 raises when the other cloud is shorter than the query cloud
        n_other = cloud_pair._n[1 if is_left else 0]
        raise IndexError(f"index {n_other} is out of bounds for axis 0 with size {n_other}")
    return fd


class GeoMSE(SecondaryMetric, PointToPlaneable):
    def _get_dependencies(self):
        return {"euclidean_distance": EuclideanDistance(is_left=self.is_left, point_to_plane=self.point_to_plane)}

    def calculate(self, euclidean_distance) -> None:
        x = euclidean_distance.value
        self.value = np.sum(x, axis=0) / x.shape[0]

    def calculate_fused(self, cloud_pair) -> bool:
        fd = _fused_dir(cloud_pair, self.is_left, self.point_to_plane)
        total = fd.sum_d2 if self.point_to_plane else fd.sum_d1
        self.value = np.float64(total) / fd.n
        return True


class GeoPSNR(SecondaryMetric, PointToPlaneable):
    def _get_dependencies(self):
        return {"cloud_extent": CloudExtent(),
                "geo_mse": GeoMSE(is_left=self.is_left, point_to_plane=self.point_to_plane)}

    def calculate(self, cloud_extent, geo_mse) -> None:
        peak = np.max(cloud_extent.value)
        with np.errstate(divide="ignore"):
            self.value = 10 * np.log10(peak ** 2 / geo_mse.value)


class GeoHausdorffDistance(SecondaryMetric, PointToPlaneable):
    def _get_dependencies(self):
        return {"euclidean_distance": EuclideanDistance(is_left=self.is_left, point_to_plane=self.point_to_plane)}

    def calculate(self, euclidean_distance) -> None:
        self.value = np.max(euclidean_distance.value, axis=0)   # max of SQUARED values (quirk Q5)

    def calculate_fused(self, cloud_pair) -> bool:
        fd = _fused_dir(cloud_pair, self.is_left, self.point_to_plane)
        self.value = np.float64(fd.max_d2 if self.point_to_plane else fd.max_d1)
        return True


class GeoHausdorffDistancePSNR(SecondaryMetric, PointToPlaneable):
    def _get_dependencies(self):
        return {"max_sqrt": MaxSqrtDistance(),
                "hausdorff_distance": GeoHausdorffDistance(is_left=self.is_left, point_to_plane=self.point_to_plane)}

    def calculate(self, max_sqrt, hausdorff_distance) -> None:
        with np.errstate(divide="ignore"):
            self.value = 10 * np.log10(max_sqrt.value ** 2 / hausdorff_distance.value)


# ---- colour (metric.py:261-350, 389-443) -----------------------------------------------------
def transform_colors(colors: np.ndarray, source_scheme: str, target_scheme: str) -> np.ndarray:
    if source_scheme == target_scheme:
        return colors
    if source_scheme != "rgb" or target_scheme not in ("ycc", "yuv"):
        raise TypeError(f"no colour transform {source_scheme} -> {target_scheme}")
    T = _COLOR_TRANSFORMS[target_scheme]
    c = np.asarray(colors, dtype=np.float64)
    return np.stack([(T[k][0] * c[:, 0] + T[k][1] * c[:, 1]) + T[k][2] * c[:, 2] for k in range(3)], axis=1)


def get_color_peak(color_scheme: str) -> float:
    return {"rgb": 255.0, "ycc": 1.0, "yuv": 1.0}[color_scheme]


def _color_diff(scheme, origin_cloud_colors, neighbour_cloud_colors):
    o = transform_colors(np.copy(origin_cloud_colors.value), "rgb", scheme)
    n = transform_colors(np.copy(neighbour_cloud_colors.value), "rgb", scheme)
    return np.subtract(o, n)


class ColorMSE(SecondaryMetric, ColorMetric):
    def _get_dependencies(self):
        return {"origin_cloud_colors": CloudColors(is_left=self.is_left),
                "neighbour_cloud_colors": NeighbourColors(is_left=self.is_left)}

    def calculate(self, origin_cloud_colors, neighbour_cloud_colors) -> None:
        diff = _color_diff(self.color_scheme, origin_cloud_colors, neighbour_cloud_colors)
        self.value = np.mean(diff ** 2, axis=0)

    def calculate_fused(self, cloud_pair) -> bool:
        fd = _fused_dir(cloud_pair, self.is_left, color_scheme=self.color_scheme)
        self.value = fd.color_sum / fd.n
        return True


class ColorPSNR(SecondaryMetric, ColorMetric):
    def _get_dependencies(self):
        return {"color_mse": ColorMSE(is_left=self.is_left, color_scheme=self.color_scheme)}

    def calculate(self, color_mse) -> None:
        peak = get_color_peak(self.color_scheme)
        with np.errstate(divide="ignore"):
            self.value = 10 * np.log10(peak ** 2 / color_mse.value)


class ColorHausdorffDistance(SecondaryMetric, ColorMetric):
    def _get_dependencies(self):
        return {"origin_cloud_colors": CloudColors(is_left=self.is_left),
                "neighbour_cloud_colors": NeighbourColors(is_left=self.is_left)}

    def calculate(self, origin_cloud_colors, neighbour_cloud_colors) -> None:
        diff = _color_diff(self.color_scheme, origin_cloud_colors, neighbour_cloud_colors)
        if self.color_scheme == "rgb":
            diff = 255 * diff                               # quirk Q6 (metric.py:421-424)
        self.value = np.max(diff ** 2, axis=0)

    def calculate_fused(self, cloud_pair) -> bool:
        fd = _fused_dir(cloud_pair, self.is_left, color_scheme=self.color_scheme)
        self.value = fd.color_max.copy()
        return True


class ColorHausdorffDistancePSNR(SecondaryMetric, ColorMetric):
    def _get_dependencies(self):
        return {"hausdorff_distance": ColorHausdorffDistance(is_left=self.is_left, color_scheme=self.color_scheme)}

    def calculate(self, hausdorff_distance) -> None:
        peak = get_color_peak(self.color_scheme)
        with np.errstate(divide="ignore"):
            self.value = 10 * np.log10(peak ** 2 / hausdorff_distance.value)


# ---- pooling (metric.py:446-485) ---------------------------------------------------------------
class SymmetricMetric(SecondaryMetric):
    def __init__(self, metrics, is_proportional: bool):
        if len(metrics) != 2:
            raise ValueError("Must be exactly two metrics")
        if type(metrics[0]) is not type(metrics[1]):
            raise ValueError(f"Metrics must be of same class, got: {type(metrics[0])}, {type(metrics[1])}")
        self.metrics = metrics
        self.is_proportional = is_proportional

    def _get_dependencies(self):
        return {"lmetric": self.metrics[0], "rmetric": self.metrics[1]}

    def _key(self) -> tuple:
        return super()._key() + self.metrics[0]._key() + self.metrics[1]._key()

    def calculate(self, lmetric, rmetric) -> None:
        pick = min if self.is_proportional else max      # PSNR-like: the worse (smaller) side
        self.value = pick([lmetric.value, rmetric.value], key=np.linalg.norm)
