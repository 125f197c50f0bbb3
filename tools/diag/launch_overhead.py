#!/usr/bin/env python
"""Fused kernel: time per launch against the number of tile rounds (592 tiles = one round on 148 SMs x 4 groups),
back-to-back launches, to separate the fixed cost per launch from the steady-state cost per round."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "neural-network-based-viterbi-decoding_b200"))
from nvd import montecarlo, weights

m = weights.load_default_fnn()
with montecarlo.Simulator(0) as sim:
    h = sim.load_mlp(m)
    c = sim.zeros(32)
    rows = []
    for rounds in (50, 100, 200, 221, 222, 400, 800):
        frames = 592 * 128 * rounds
        for _ in range(3):
            sim.run_async(h, "133171", 7, [2.0], frames, c, 0, mode="fused")
        sim.sync()
        e0, e1 = sim.event(), sim.event()
        sim.record(e0)
        for s in range(20):
            sim.run_async(h, "133171", 7, [2.0], frames, c, s * frames, mode="fused")
        sim.record(e1); sim.sync()
        ms = sim.elapsed_ms(e0, e1) / 20
        rows.append((rounds, ms))
        print(f"rounds {rounds:4d}  frames {frames:9d}  {ms*1e3:9.2f} us/launch  {ms*1e3/rounds:7.4f} us/round  {frames*7/ms/1e6:7.2f} Gbit/s")
    r = np.array([x[0] for x in rows], float); t = np.array([x[1] for x in rows], float) * 1e3
    b, a = np.polyfit(r, t, 1)
    print(f"fit: {a:.2f} us fixed per launch + {b:.4f} us per round ({b*1965/4:.0f} clk/tile at 1965 MHz)")
    frames = 1 << 24
    print(f"2^24 frames = {frames/128/592:.2f} rounds: predicted {a + b*np.ceil(frames/128/592):.1f} us (whole rounds), {a + b*frames/128/592:.1f} us (fractional)")
