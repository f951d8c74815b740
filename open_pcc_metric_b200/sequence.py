"""Sequence / batch evaluation (BASELINE config 4: frames of a
dynamic cloud, sharded over the GPUs of one box).

The reference cannot do this at all: its memo is a class attribute (calculator.py:60, quirk
Q2), so a second pair in the same process returns the first pair's values.  Here every frame
gets its own CloudPair and MetricCalculator; frame t runs on rank t mod world (no collective on
the data path) and rank 0 gathers the per-frame rows at the end.

Within a rank the frames are PIPELINED two deep (``pipeline=2``, the default) over two contexts (= two sets
of CUDA streams): building a CloudPair only enqueues work, so frame t+1 is loaded, narrowed and uploaded
while the kernels of frame t run, and frame t's one host wait (its result read-back) overlaps frame t+1's
copies.  Rows are delivered -- and, with ``csv_path``, appended to the CSV file -- in frame order.
"""
from __future__ import annotations

import typing

import pandas as pd

from . import _native as N
from .calculator import MetricCalculator
from .cloud_pair import CloudPair
from .geometry import default_context
from .options import CalculateOptions, transform_options

_COLUMNS = ["frame", "label", "is_left", "point-to-plane", "value"]


def _evaluate_frame(t, item, options, ctx, pair_kwargs):
    a, b = item() if callable(item) else item
    pair = CloudPair(a, b, ctx=ctx, **pair_kwargs)
    try:
        df = MetricCalculator(pair).calculate(transform_options(options)).as_df()
    finally:
        pair.close()
    df.insert(0, "frame", t)
    return df


def _finish_frame(t, pair, options):
    try:
        df = MetricCalculator(pair).calculate(transform_options(options)).as_df()
    finally:
        pair.close()
    df.insert(0, "frame", t)
    return df


def evaluate_sequence(frames: typing.Iterable[typing.Tuple[typing.Any, typing.Any]],
                      options: CalculateOptions | None = None, *, ctx: N.Context | None = None,
                      rank: int = 0, world: int = 1, group=None, pipeline: int = 2,
                      csv_path: str | None = None, **pair_kwargs) -> pd.DataFrame | None:
    """frames: iterable of (origin_cloud, reconst_cloud); items may be callables returning the
    pair (lazy loading: only the frames of this rank are materialised, inside the worker that
    evaluates them).  Returns the reference's result table (calculator.py:27-52) with a leading
    ``frame`` column on rank 0 (None elsewhere when world > 1).  csv_path: every finished frame's rows
    are appended to ``csv_path`` (rank r > 0 writes ``csv_path + ".rank<r>"``) as soon as all earlier
    frames of the rank have been written."""
    options = options or CalculateOptions()
    ctx = ctx or default_context()
    mine = [(t, item) for t, item in enumerate(frames) if t % world == rank]
    tables: list[pd.DataFrame] = []
    path = None if csv_path is None else (csv_path if rank == 0 else f"{csv_path}.rank{rank}")
    wrote_header = False

    def deliver(df):
        nonlocal wrote_header
        tables.append(df)
        if path is not None:
            df.to_csv(path, mode="a" if wrote_header else "w", header=not wrote_header, index=False)
            wrote_header = True

    if int(pipeline) <= 1 or len(mine) < 2:
        for t, item in mine:
            deliver(_evaluate_frame(t, item, options, ctx, pair_kwargs))
    else:
        # software pipeline over two contexts (= two sets of CUDA streams): constructing a CloudPair only ENQUEUES its
        # uploads, statistics and index build, so frame t+1 is loaded, narrowed and sent while the kernels of frame t run;
        # frame t is then evaluated (its one host wait) while frame t+1's copies are still in flight
        ctxs = [ctx, N.Context(ctx.device)]
        try:
            prev = None
            for k, (t, item) in enumerate(mine):
                a, b = item() if callable(item) else item
                pair = CloudPair(a, b, ctx=ctxs[k % 2], **pair_kwargs)
                if prev is not None:
                    deliver(_finish_frame(*prev, options))
                prev = (t, pair)
            if prev is not None:
                deliver(_finish_frame(*prev, options))
        finally:
            ctxs[1].close()
    local_df = pd.concat(tables, ignore_index=True) if tables else pd.DataFrame(columns=_COLUMNS)
    if world == 1:
        return local_df
    import torch.distributed as dist
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(local_df, gathered, dst=0, group=group)
    if rank != 0:
        return None
    out = pd.concat(gathered, ignore_index=True)
    return out.sort_values(["frame"], kind="stable").reset_index(drop=True)
