#!/usr/bin/env python
"""K4 throughput: persistent LSTM seq2seq decoder, (133,171) L = 100 blocks (BASELINE.json configs[3])."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "neural-network-based-viterbi-decoding_b200"))
from nvd import montecarlo, weights
from oracle import lstm as olstm

L, H = 100, 20
B = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 128 * 10   # two rounds of 5 groups per SM
p = olstm.init_params(H, seed=4)
with montecarlo.Simulator(0) as sim:
    h = sim.load_lstm(weights.LstmModel(**p))
    rng = np.random.default_rng(0)
    y = rng.standard_normal((B, 2 * L)).astype(np.float32)
    d_y = sim.to_device(y)
    from nvd import _lib
    d_hard = sim.alloc(B * 25)            # buffers allocated once: no cudaMalloc inside the timed region
    def step(fast):
        _lib.check(sim.lib.nvd_lstm_forward(sim.ctx, h, d_y.ptr, _lib.DTYPES["f32"], None, 2 * L, 2 * L, B, fast, None, d_hard.ptr))
    for fast in (0, 1, 2):
        for _ in range(2):
            step(fast)
        sim.sync()
        e0, e1 = sim.event(), sim.event()
        sim.record(e0)
        reps = 5
        for _ in range(reps):
            step(fast)
        sim.record(e1); sim.sync()
        ms = sim.elapsed_ms(e0, e1) / reps
        fma = B * 400 * (1 + H) * H * 4
        print(f"fast_math={fast}: B={B} {ms:.3f} ms  {B / ms / 1e3:.2f} Mframes/s  {B * L / ms / 1e6:.3f} Gbit/s  "
              f"{fma / ms / 1e9:.2f} TFMA/s ({fma / ms / 1e9 / (148 * 128 * 1.965e-3):.1%} of the FP32 FMA peak)")
