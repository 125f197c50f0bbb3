#!/usr/bin/env python
"""Wall-clock time of the deep-BER sweep (nvd_mc_sweep: a host decision after every chunk) — isolated launches."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "neural-network-based-viterbi-decoding_b200"))
from nvd import montecarlo, weights
with montecarlo.Simulator(0) as sim:
    h = sim.load_mlp(weights.load_default_fnn())
    c = sim.zeros(32)
    for rep in range(4):
        for _ in range(10):
            sim.run_async(h, "133171", 7, [2.0], 1 << 24, c, 0, mode="fused")
        sim.sync()
        t0 = time.perf_counter()
        dc = sim.sweep(h, "133171", 7, [5.0, 6.0, 7.0, 8.0], min_frame_errors=100, max_frames=1 << 34, chunk_frames=1 << 26, mode="fused", world=1)
        dt = time.perf_counter() - t0
        fr = int(dc[:, 2].sum())
        print(f"sweep: {fr} frames in {dt*1e3:.2f} ms -> {fr*7/dt/1e9:.1f} Gbit/s; frame errors {dc[:,1].tolist()}")
