"""The `utils` package production_wrapper.py imports (config, logger, metrics, checkpoint) — absent
from the reference tree (production_wrapper.py:26-29); its API is fixed by the wrapper's call sites."""
