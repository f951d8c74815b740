"""Sequence / batch evaluation (BASELINE config 4: frames of a
dynamic cloud, sharded over the GPUs of one box).

The reference cannot do this at all: its memo is a class attribute (calculator.py:60, quirk
Q2), so a second pair in the same process returns the first pair's values.  Here every frame
gets its own CloudPair and MetricCalculator; frame t runs on rank t mod world (no collective on
the data path) and rank 0 gathers the per-frame rows at the end.
"""
from __future__ import annotations

import typing

import pandas as pd

from . import _native as N
from .calculator import MetricCalculator
from .cloud_pair import CloudPair
from .geometry import default_context
from .options import CalculateOptions, transform_options


def evaluate_sequence(frames: typing.Iterable[typing.Tuple[typing.Any, typing.Any]],
                      options: CalculateOptions | None = None, *, ctx: N.Context | None = None,
                      rank: int = 0, world: int = 1, group=None, **pair_kwargs) -> pd.DataFrame | None:
    """frames: iterable of (origin_cloud, reconst_cloud); items may be callables returning the
    pair (lazy loading: only the frames of this rank are materialised).  Returns the reference's
    result table (calculator.py:27-52) with a leading ``frame`` column on rank 0 (None elsewhere
    when world > 1)."""
    options = options or CalculateOptions()
    ctx = ctx or default_context()
    tables = []
    for t, item in enumerate(frames):
        if t % world != rank:
            continue
        a, b = item() if callable(item) else item
        pair = CloudPair(a, b, ctx=ctx, **pair_kwargs)
        df = MetricCalculator(pair).calculate(transform_options(options)).as_df()
        df.insert(0, "frame", t)
        tables.append(df)
        pair.close()
    local = pd.concat(tables, ignore_index=True) if tables else pd.DataFrame(
        columns=["frame", "label", "is_left", "point-to-plane", "value"])
    if world == 1:
        return local
    import torch.distributed as dist
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(local, gathered, dst=0, group=group)
    if rank != 0:
        return None
    out = pd.concat(gathered, ignore_index=True)
    return out.sort_values(["frame"], kind="stable").reset_index(drop=True)
