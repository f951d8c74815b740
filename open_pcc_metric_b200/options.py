"""CalculateOptions -> ordered metric list, same content and order as the reference's
transform_options (/root/reference/open_pcc_metric/options.py:32-174): 8 base
entries, then +6 colour, +6 point-to-plane, +6 Hausdorff, +6 Hausdorff x plane."""
from __future__ import annotations

import typing

from .metric import (AbstractMetric, ColorMSE, ColorPSNR, GeoHausdorffDistance, GeoHausdorffDistancePSNR,
                     GeoMSE, GeoPSNR, MaxSqrtDistance, MinSqrtDistance, SymmetricMetric)


class CalculateOptions:
    def __init__(self, color: typing.Optional[str] = None, hausdorff: bool = False, point_to_plane: bool = False):
        self.color = color
        self.hausdorff = hausdorff
        self.point_to_plane = point_to_plane


def _both(cls, **kw):
    return [cls(is_left=True, **kw), cls(is_left=False, **kw)]


def _pooled(cls, proportional, **kw):
    return SymmetricMetric(metrics=tuple(_both(cls, **kw)), is_proportional=proportional)


def _block(error_cls, psnr_cls, **kw):
    """left, right, pooled error (max) then left, right, pooled PSNR (min)."""
    return _both(error_cls, **kw) + [_pooled(error_cls, False, **kw)] + _both(psnr_cls, **kw) + [_pooled(psnr_cls, True, **kw)]


def transform_options(options: CalculateOptions) -> typing.List[AbstractMetric]:
    metrics: typing.List[AbstractMetric] = [MinSqrtDistance(), MaxSqrtDistance()]
    metrics += _block(GeoMSE, GeoPSNR, point_to_plane=False)
    if options.color is not None:
        metrics += _block(ColorMSE, ColorPSNR, color_scheme=options.color)
    if options.point_to_plane:
        metrics += _block(GeoMSE, GeoPSNR, point_to_plane=True)
    if options.hausdorff:
        metrics += _block(GeoHausdorffDistance, GeoHausdorffDistancePSNR, point_to_plane=False)
    if options.hausdorff and options.point_to_plane:
        # options.py:140-172 lists this block in a different order: l, r, l-psnr, r-psnr, pooled, pooled-psnr
        kw = dict(point_to_plane=True)
        metrics += _both(GeoHausdorffDistance, **kw) + _both(GeoHausdorffDistancePSNR, **kw)
        metrics += [_pooled(GeoHausdorffDistance, False, **kw), _pooled(GeoHausdorffDistancePSNR, True, **kw)]
    return metrics
