"""setup_logger(name, log_level, log_dir, use_json, log_to_console, log_to_file) — production_wrapper.py:51-58."""
from __future__ import annotations

import json
import logging
import os
import time


class _JsonFormatter(logging.Formatter):
    def format(self, record):
        return json.dumps({"ts": time.strftime("%Y-%m-%dT%H:%M:%S", time.gmtime(record.created)), "level": record.levelname,
                           "logger": record.name, "message": record.getMessage()})


def setup_logger(name: str = "neural_viterbi", log_level: str = "INFO", log_dir: str = "./logs", use_json: bool = False,
                 log_to_console: bool = True, log_to_file: bool = True) -> logging.Logger:
    logger = logging.getLogger(name)
    logger.setLevel(getattr(logging, str(log_level).upper(), logging.INFO))
    logger.handlers.clear()
    fmt = _JsonFormatter() if use_json else logging.Formatter("%(asctime)s %(levelname)s %(name)s: %(message)s")
    if log_to_console:
        h = logging.StreamHandler()
        h.setFormatter(fmt)
        logger.addHandler(h)
    if log_to_file:
        os.makedirs(log_dir, exist_ok=True)
        h = logging.FileHandler(os.path.join(log_dir, f"{name}.log"))
        h.setFormatter(fmt)
        logger.addHandler(h)
    logger.propagate = False
    return logger
