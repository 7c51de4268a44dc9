"""Rank-0 stdout logger (same contract as the reference's
video_to_video/utils/logger.py:14-69: one shared logger named after the top
package, INFO on rank 0, ERROR elsewhere, no propagation to the root logger)."""
import logging

import torch.distributed as dist

_configured = {}
_FORMAT = logging.Formatter("%(asctime)s - %(name)s - %(levelname)s - %(message)s")


def _is_master():
    return not (dist.is_available() and dist.is_initialized()) or dist.get_rank() == 0


def get_logger(log_file=None, log_level=logging.INFO, file_mode="w"):
    name = __name__.split(".")[0]
    logger = logging.getLogger(name)
    logger.propagate = False
    if name not in _configured:
        h = logging.StreamHandler()
        h.setFormatter(_FORMAT)
        logger.addHandler(h)
        _configured[name] = True
    if log_file is not None and _is_master() and not any(
            isinstance(h, logging.FileHandler) for h in logger.handlers):
        fh = logging.FileHandler(log_file, file_mode)
        fh.setFormatter(_FORMAT)
        logger.addHandler(fh)
    level = log_level if _is_master() else logging.ERROR
    logger.setLevel(level)
    for h in logger.handlers:
        h.setLevel(level)
    return logger
