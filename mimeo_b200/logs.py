"""Logging setup (host mirror of mimeo.logs.init_logging, logs.py:15-54; stdlib handler instead of rich)."""
import logging
import sys


def init_logging(loglevel: str = 'DEBUG') -> None:
    numeric = getattr(logging, str(loglevel).upper(), None)
    if not isinstance(numeric, int):
        raise ValueError(f'Invalid log level: {loglevel}')
    handler = logging.StreamHandler(sys.stderr)
    handler.setFormatter(logging.Formatter('%(asctime)s | %(levelname)s | %(message)s', '%Y-%m-%d %H:%M:%S'))
    logging.basicConfig(level=numeric, handlers=[handler], force=True)
