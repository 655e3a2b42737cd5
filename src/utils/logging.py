"""`get_logger` with the reference's signature (`src/utils/logging.py:10-13`)."""
import logging


def get_logger(name: str):
    logger = logging.getLogger(name)
    logger.setLevel(logging.INFO)
    return logger


# names this file does not provide resolve from a reference checkout's copy of the same file (see src/__init__.py)
import src as _src  # noqa: E402

__getattr__ = _src.fallback_getattr(__name__)
