"""Logger accessor (reference: logger.py:5-16).  Unlike the reference it does not
stack a new stderr handler on the root logger at every call (quirk Q17)."""
import logging
import sys

_NAME = "open_pcc_metric_b200"


def get_logger() -> logging.Logger:
    log = logging.getLogger(_NAME)
    if not log.handlers:
        h = logging.StreamHandler(sys.stderr)
        h.setFormatter(logging.Formatter("%(asctime)s - %(name)s - %(levelname)s - %(message)s"))
        log.addHandler(h)
        log.setLevel(logging.WARNING)
    return log
