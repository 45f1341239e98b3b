"""``pynext`` logger, the name the reference's gravity module logs to (pynbodyext/log.py:4)."""
import logging

logger = logging.getLogger("pynext")
if not logger.handlers:
    logger.addHandler(logging.NullHandler())
