"""Minimal stand-in for the reference's ngen-style logging module (reference ``bmi/logger.py``).

Only the two names the reference package exports are provided (``logger``, ``configure_logging``); the
environment variables ``NGEN_EWTS_LOGGING`` / ``TOPOFLOW_GLACIER_LOGLEVEL`` are honoured.  Logging is
observability, not part of the numeric path.
"""

from __future__ import annotations

import logging
import os

logger = logging.getLogger("Topoflow-Glacier")
_configured = False


def configure_logging() -> None:
    global _configured
    if _configured:
        return
    _configured = True
    if os.environ.get("NGEN_EWTS_LOGGING", "ENABLED").upper() == "DISABLED":
        logger.addHandler(logging.NullHandler())
        logger.propagate = False
        return
    level = os.environ.get("TOPOFLOW_GLACIER_LOGLEVEL", "WARNING").upper()
    level = {"SEVERE": "ERROR", "FATAL": "CRITICAL"}.get(level, level)
    logger.setLevel(getattr(logging, level, logging.WARNING))
    if not logger.handlers:
        h = logging.StreamHandler()
        h.setFormatter(logging.Formatter("%(asctime)s %(name)s %(levelname)s %(message)s"))
        logger.addHandler(h)
