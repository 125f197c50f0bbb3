#!/usr/bin/env python
"""Sliding-window decoder throughput (fnn_viterbi_seq_predict, 200-500x4-2 tanh MLP over stride-2 windows of one
stream): fp32 layer kernel vs the wide-layer tcgen05 kernel."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "neural-network-based-viterbi-decoding_b200"))
from nvd import montecarlo, weights, _lib
from oracle import mlp as omlp

n_win = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
dims, acts = [200, 500, 500, 500, 500, 2], [1, 1, 1, 1, 1]
Ws, bs = omlp.init_fnn(dims, seed=3, bias_scale=0.05)
mac = sum(a * b for a, b in zip(dims[:-1], dims[1:]))
with montecarlo.Simulator(0) as sim:
    h = sim.load_mlp(weights.MlpModel(Ws, bs, acts))
    n_sym = 2 * (n_win - 1) + 200
    d_stream = sim.to_device(np.random.default_rng(0).standard_normal(n_sym).astype(np.float32))
    d_hard = sim.alloc(n_win)
    for prec, reps, nw in (("fp32", 1, min(n_win, 1 << 16)), ("bf16", 5, n_win)):
        def step():
            _lib.check(sim.lib.nvd_window_decode(sim.ctx, h, d_stream.ptr, _lib.DTYPES["f32"], n_sym, 2, nw, _lib.PRECISIONS[prec], None, d_hard.ptr, 1))
        step(); sim.sync()
        e0, e1 = sim.event(), sim.event()
        sim.record(e0)
        for _ in range(reps):
            step()
        sim.record(e1); sim.sync()
        ms = sim.elapsed_ms(e0, e1) / reps
        print(f"{prec}: {nw} windows in {ms:.3f} ms  {nw / ms / 1e3:.2f} Mbit/s decoded  {2 * mac * nw / ms / 1e9:.1f} TFLOP/s "
              f"({2 * mac * nw / ms / 1e9 / 1656.9:.1%} of the measured bf16 burst peak)")
