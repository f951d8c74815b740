"""MetricCalculator / CalculateResult -- mirrors the reference's calculator.py
(/root/reference/open_pcc_metric/calculator.py:15-108): recursive, memoised
resolution of the metric graph and the result table.

Deviations, both deliberate:
* the memo is PER INSTANCE.  In the reference it is a class attribute
  (calculator.py:60, quirk Q2), so a second calculator in one process returns the
  first pair's values; sequence evaluation needs the per-instance form.
* reductions that the GPU already produced are taken from ``calculate_fused``
  before the dependency graph is walked (``use_fused=False`` restores the pure
  graph walk over per-point arrays).
"""
from __future__ import annotations

import typing

import pandas as pd

from .cloud_pair import CloudPair
from .metric import AbstractMetric, PrimaryMetric, SecondaryMetric, SymmetricMetric


class CalculateResult:
    def __init__(self, metrics: typing.List[AbstractMetric]):
        self._metrics = metrics

    def as_dict(self) -> typing.Dict[tuple, typing.Any]:
        return {m._key(): m.value for m in self._metrics}

    def as_df(self) -> pd.DataFrame:
        rows = {"label": [], "is_left": [], "point-to-plane": [], "value": []}
        for m in self._metrics:
            label = type(m).__name__
            if isinstance(m, SymmetricMetric):
                label = type(m.metrics[0]).__name__ + "(symmetric)"
            rows["label"].append(label)
            rows["is_left"].append(getattr(m, "is_left", ""))
            rows["point-to-plane"].append(getattr(m, "point_to_plane", ""))
            rows["value"].append(str(m.value))
        return pd.DataFrame(rows)

    def __str__(self) -> str:
        return str(self.as_df())


class MetricCalculator:
    def __init__(self, cloud_pair: CloudPair, use_fused: bool = True):
        self._cloud_pair = cloud_pair
        self._calculated_metrics: typing.Dict[tuple, AbstractMetric] = {}
        self._use_fused = use_fused and hasattr(cloud_pair, "fused")

    def _metric_recursive_calculate(self, metric: AbstractMetric) -> AbstractMetric:
        key = metric._key()
        done = self._calculated_metrics.get(key)
        if done is not None:
            return done
        if isinstance(metric, PrimaryMetric):
            metric.calculate(self._cloud_pair)
        elif isinstance(metric, SecondaryMetric):
            fused = getattr(metric, "calculate_fused", None) if self._use_fused else None
            if fused is None or not fused(self._cloud_pair):
                deps = {name: self._metric_recursive_calculate(dep)
                        for name, dep in metric._get_dependencies().items()}
                metric.calculate(**deps)
        else:
            raise RuntimeError(f"Metric of unknown AbstractMetric subclass {type(metric).__name__}")
        self._calculated_metrics[key] = metric
        return metric

    def _prefetch(self, metrics_list) -> None:
        """One fused GPU evaluation for everything the list will ask for (D2 and the first colour
        scheme on top of D1) instead of one per kind of metric as the graph walk reaches them."""
        p2p, schemes, boundary = False, [], False

        def scan(m):
            nonlocal p2p, boundary
            if type(m).__name__ in ("MinSqrtDistance", "MaxSqrtDistance", "BoundarySqrtDistances", "GeoHausdorffDistancePSNR"):
                boundary = True
            if isinstance(m, SymmetricMetric):
                for c in m.metrics:
                    scan(c)
                return
            if not hasattr(m, "calculate_fused") and not hasattr(m, "_get_dependencies"):
                return
            if getattr(m, "point_to_plane", False):
                p2p = True
            cs = getattr(m, "color_scheme", None)
            if cs is not None and cs not in schemes:
                schemes.append(cs)
        for m in metrics_list:
            scan(m)
        if boundary and hasattr(self._cloud_pair, "boundary_minmax"):
            # needs coordinates only: runs while colours / normals are still on their way to the device
            try:
                self._cloud_pair.boundary_minmax()
            except (IndexError, ValueError, KeyError):
                pass
        if p2p or schemes:
            try:
                self._cloud_pair.fused(True, point_to_plane=p2p, color_scheme=schemes[0] if schemes else None)
            except (IndexError, ValueError, KeyError):
                pass    # the individual metric raises the reference's error when it is reached

    def calculate(self, metrics_list: typing.List[AbstractMetric]) -> CalculateResult:
        if self._use_fused:
            self._prefetch(metrics_list)
        return CalculateResult([self._metric_recursive_calculate(m) for m in metrics_list])
