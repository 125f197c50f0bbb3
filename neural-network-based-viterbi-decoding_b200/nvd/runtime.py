"""Process-wide runtime used by the drop-in modules (src/*.py): one lazily created Simulator per
process and a global frame counter so that successive `modulate_awgn` calls draw fresh,
reproducible Philox noise (the reference draws from numpy's unseeded global RNG,
/root/reference/src/fnn_viterbi.py:250).  Creating the Simulator fails loudly without a GPU."""
from __future__ import annotations

import os
import threading

from .montecarlo import DEFAULT_SEED, Simulator

_lock = threading.Lock()
_sim: Simulator | None = None
_next_frame = 0
_precision = os.environ.get("NVD_PRECISION", "fp32")


def get_sim() -> Simulator:
    global _sim
    with _lock:
        if _sim is None:
            device = int(os.environ.get("NVD_DEVICE", os.environ.get("LOCAL_RANK", "0")))
            _sim = Simulator(device=device, seed=int(os.environ.get("NVD_SEED", DEFAULT_SEED)))
        return _sim


def seed(value: int) -> None:
    """Re-seed the channel noise and restart the frame counter (reproducible data sets)."""
    global _next_frame
    get_sim().set_seed(int(value))
    _next_frame = 0


def take_frames(n: int) -> int:
    """Reserve n global frame indices; returns the first."""
    global _next_frame
    with _lock:
        f0 = _next_frame
        _next_frame += int(n)
        return f0


def set_precision(p: str) -> None:
    """'fp32' (FFMA parity path, default) or 'bf16' (tcgen05 tensor-core path where the shape is covered)."""
    global _precision
    if p not in ("fp32", "bf16"):
        raise ValueError(p)
    _precision = p


def precision() -> str:
    return _precision


def shutdown() -> None:
    global _sim
    with _lock:
        if _sim is not None:
            _sim.close()
            _sim = None
