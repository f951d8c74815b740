"""open_pcc_metric_b200 -- B200 (sm_100a) build of open-pcc-metric's hot path.

Drop-in surface (same names as /root/reference/open_pcc_metric): ``CloudPair``,
the ``metric`` classes, ``MetricCalculator``, ``CalculateOptions`` /
``transform_options`` and the ``cli``.  All geometry runs in libpccm.so
(include/pccm.h); there is no CPU fallback.
"""
from . import _native  # noqa: F401  (does not load the library until first use)

__all__ = ["CloudPair", "MetricCalculator", "CalculateOptions", "transform_options", "PointCloud"]
__version__ = "0.1.0"


def __getattr__(name):  # lazy: importing the package must not need pandas / click / a GPU
    if name == "CloudPair":
        from .cloud_pair import CloudPair
        return CloudPair
    if name == "MetricCalculator":
        from .calculator import MetricCalculator
        return MetricCalculator
    if name in ("CalculateOptions", "transform_options"):
        from . import options
        return getattr(options, name)
    if name == "PointCloud":
        from .geometry import PointCloud
        return PointCloud
    raise AttributeError(name)
