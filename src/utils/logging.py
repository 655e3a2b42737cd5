"""`get_logger` with the reference's signature (`src/utils/logging.py:10-13`)."""
import logging


def get_logger(name: str):
    logger = logging.getLogger(name)
    logger.setLevel(logging.INFO)
    return logger
