"""``python -m open_pcc_metric_b200 --ocloud a.ply --pcloud b.ply [...]``: runs the click
command of handler.py (same flags and table as the reference CLI, on a B200)."""
import sys

from .handler import cli

sys.exit(cli())  # pylint: disable=no-value-for-parameter
