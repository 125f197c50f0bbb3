#!/usr/bin/env python
"""Fast iteration loop for the fused kernel: ms/step at 2^24 frames + fused==staged check."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "neural-network-based-viterbi-decoding_b200"))
from nvd import montecarlo, weights

frames = 1 << 24
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
m = weights.load_default_fnn()
with montecarlo.Simulator(0) as sim:
    h = sim.load_mlp(m)
    c = sim.zeros(32)
    for _ in range(3):
        sim.run_async(h, "133171", 7, [2.0], frames, c, 0, mode="fused")
    sim.sync()
    e0, e1 = sim.event(), sim.event()
    sim.record(e0)
    for s in range(steps):
        sim.run_async(h, "133171", 7, [2.0], frames, c, s * frames, mode="fused")
    sim.record(e1); sim.sync()
    ms = sim.elapsed_ms(e0, e1) / steps
    print(f"fused: {ms:.4f} ms/step  {frames * 7 / ms / 1e6:.2f} Gbit/s  {frames / ms / 1e6:.3f} Gframes/s  "
          f"{frames/128/148/ (ms*1e-3) :.0f} tiles/s/SM -> {1.965e9/(frames/128/148/(ms*1e-3)):.0f} clk/tile")
    a = sim.run(h, "133171", 7, [2.0, 5.0], 1 << 20, 12345, mode="fused")
    b = sim.run(h, "133171", 7, [2.0, 5.0], 1 << 20, 12345, mode="staged", precision="bf16")
    print("fused", a.tolist(), "staged", b.tolist(), "EQUAL" if (a == b).all() else "MISMATCH")
    e0, e1 = sim.event(), sim.event()
    c2 = sim.zeros(32)
    sim.run_async(h, "133171", 7, [2.0], frames, c2, 0, mode="staged", precision="bf16")
    sim.sync(); sim.record(e0)
    for s in range(5):
        sim.run_async(h, "133171", 7, [2.0], frames, c2, 0, mode="staged", precision="bf16")
    sim.record(e1); sim.sync()
    print(f"staged bf16: {sim.elapsed_ms(e0, e1)/5:.4f} ms/step")
