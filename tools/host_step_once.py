#!/usr/bin/env python
"""A few end-to-end host-buffer steps (nvd_mc_run_host, streaming form when NVD_HOST_STEP=streaming) — the command the
MODE_FUSED_IO ncu capture profiles with --replay-mode application."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "neural-network-based-viterbi-decoding_b200"))
from nvd import montecarlo, weights
frames = 1 << 22
with montecarlo.Simulator(0) as sim:
    h = sim.load_mlp(weights.load_default_fnn())
    info, pred = sim.pinned(frames), sim.pinned(frames)
    info[:] = montecarlo.class_to_packed(np.arange(128), 7).astype(np.uint8)[np.arange(frames) & 127]
    for _ in range(3):
        c = sim.run_host(h, "133171", 7, 2.0, info, pred, mode="fused")
    print("counters", c.tolist())
