#!/usr/bin/env python
"""Does an idle gap before the timed region slow the first launches?  Per-launch CUDA-event times of 20 fused launches,
(a) right behind the warm-up launches, (b) after a synchronize + 5 ms of host sleep."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "neural-network-based-viterbi-decoding_b200"))
from nvd import montecarlo, weights
frames = 1 << 24
with montecarlo.Simulator(0) as sim:
    h = sim.load_mlp(weights.load_default_fnn())
    c = sim.zeros(32)
    ev = [sim.event() for _ in range(21)]
    for gap in (False, True, False, True):
        for _ in range(5):
            sim.run_async(h, "133171", 7, [2.0], frames, c, 0, mode="fused")
        if gap:
            sim.sync(); time.sleep(0.005)
        sim.record(ev[0])
        for s in range(20):
            sim.run_async(h, "133171", 7, [2.0], frames, c, s * frames, mode="fused")
            sim.record(ev[s + 1])
        sim.sync()
        ms = [sim.elapsed_ms(ev[i], ev[i + 1]) for i in range(20)]
        print(f"gap={gap}: total {sum(ms) / 20:.4f} ms/step; per launch:", " ".join(f"{m:.3f}" for m in ms))
