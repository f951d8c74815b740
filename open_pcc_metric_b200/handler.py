"""Command line front end -- the reference's flags (handler.py:5-43:
--ocloud --pcloud --color --hausdorff --point-to-plane --csv) with the same table
output, plus additive flags for the GPU build."""
import click


@click.command()
@click.option("--ocloud", required=True, type=str, help="Original point cloud.")
@click.option("--pcloud", required=True, type=str, help="Processed point cloud.")
@click.option("--color", required=False, type=click.Choice(["rgb", "ycc", "yuv"]),
              help="Report color distortions as well.")
@click.option("--hausdorff", required=False, is_flag=True,
              help="Report hausdorff metric as well. With --point-to-plane the point-to-plane hausdorff is reported too.")
@click.option("--point-to-plane", required=False, is_flag=True, help="Report point-to-plane distance as well.")
@click.option("--csv", required=False, is_flag=True, help="Print output in csv format.")
@click.option("--device", type=int, default=0, show_default=True, help="CUDA device ordinal.")
@click.option("--peak", type=click.Choice(["obb", "aabb_diag", "resolution"]), default="obb", show_default=True,
              help="PSNR peak: minimal oriented box extent (reference), bounding-box diagonal or 2^bits-1.")
@click.option("--bits", type=int, default=None, help="Voxel bit depth for --peak resolution.")
@click.option("--normals", "normals_mode", type=click.Choice(["reference", "neighbour"]), default="reference",
              show_default=True, help="Normal used by point-to-plane: at the query index (reference) or of the match.")
@click.option("--timings", is_flag=True, help="Print per-stage device timings (JSON) to stderr.")
@click.option("--native-dtypes", is_flag=True,
              help="Upload the PLY files' own scalar types (ushort / float coordinates, uchar colours, float normals) instead of "
                   "float64 copies: a third of the bytes over PCIe, identical metrics (every widening on the device is exact).")
def cli(ocloud, pcloud, color, hausdorff, point_to_plane, csv, device, peak, bits, normals_mode, timings, native_dtypes):
    import json
    import sys

    from . import _native as N
    from .calculator import MetricCalculator
    from .cloud_pair import CloudPair
    from .io import read_point_cloud
    from .options import CalculateOptions, transform_options

    ctx = N.Context(device)
    ctx.set_profiling(2 if timings else 0)
    # (pageable arrays: for ONE pair page-locking costs more than the staged copy it saves -- profiles/README.md)
    clouds = [read_point_cloud(p, native=native_dtypes) for p in (ocloud, pcloud)]
    pair = CloudPair(clouds[0], clouds[1], ctx=ctx, peak=peak, resolution_bits=bits, normals_mode=normals_mode)
    metrics = transform_options(CalculateOptions(color=color, hausdorff=hausdorff, point_to_plane=point_to_plane))
    table = MetricCalculator(pair).calculate(metrics).as_df()
    print(table.to_csv() if csv else table.to_string())
    if timings:
        print(json.dumps(ctx.timings()), file=sys.stderr)
