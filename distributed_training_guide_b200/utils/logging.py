"""Logging setup: ``[rank=R] [time] LEVEL:message`` (reference ``02-...:43-46``; chapter 01
has no rank prefix, ``01-single-gpu/train_llm.py:29-32``)."""
from __future__ import annotations

import logging
import os

LOGGER = logging.getLogger("dtg_b200")


def setup_logging(rank=None, level=logging.INFO):
    prefix = f"[rank={rank}] " if rank is not None else ""
    logging.basicConfig(format=f"{prefix}[%(asctime)s] %(levelname)s:%(message)s", level=level, force=True)
    LOGGER.debug(dict(os.environ))
    return LOGGER
