#!/usr/bin/env python
"""Timing of the K-streamed kernel (per-bit predictor 1036-30-2): staged rows from HBM and the fused Monte-Carlo form."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "neural-network-based-viterbi-decoding_b200"))
from nvd import montecarlo as mc, weights

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
m = weights.load_default_bitwise()
with mc.Simulator(0) as sim:
    h = sim.load_mlp(m)
    B = 1 << 19
    rng = np.random.default_rng(0)
    y16 = mc.f32_to_bf16(np.trunc(2 * rng.standard_normal((B, 1036), dtype=np.float32)))
    d_y = sim.to_device(y16)
    hard = sim.alloc(B)
    from nvd import _lib
    def step():
        _lib.check(sim.lib.nvd_mlp_forward(sim.ctx, h, d_y.ptr, _lib.DTYPES["bf16"], B, _lib.PRECISIONS["bf16"], None, hard.ptr, 1))
    step(); sim.sync()
    e0, e1 = sim.event(), sim.event()
    sim.record(e0)
    for _ in range(reps):
        step()
    sim.record(e1); sim.sync()
    ms = sim.elapsed_ms(e0, e1) / reps
    print(f"kstream staged: {ms:.4f} ms for {B} rows  {B * 2073 / ms / 1e6:.1f} GB/s  {B / ms / 1e3:.1f} Mrows/s")
    F = 1 << 21
    c = sim.zeros(32)
    sim.run_bitwise_async(h, 512, [2.0], F, c, 0, mode="fused"); sim.sync()
    sim.record(e0)
    for r in range(reps):
        sim.run_bitwise_async(h, 512, [2.0], F, c, r * F, mode="fused")
    sim.record(e1); sim.sync()
    ms = sim.elapsed_ms(e0, e1) / reps
    print(f"kstream fused : {ms:.4f} ms for {F} frames  {F / ms / 1e3:.1f} Mframes/s  {F * 1036 / ms / 1e6:.1f} Gnormals/s")
    # staged Monte-Carlo step: K2 writes rows padded to 1040 symbols, the K-streamed kernel takes them by TMA
    Fs = 1 << 19
    sim.set_stage_timing(True)
    cs = sim.zeros(32)
    for _ in range(2):
        sim.run_bitwise_async(h, 512, [2.0], Fs, cs, 0, mode="staged", trunc=True)
    ms = sim.stage_ms()
    sim.set_stage_timing(False)
    algb = {"source": 64 + 64 + 65, "encode": 65 + 130, "channel": 130 + 2072, "decode": 2072 + 1, "count": 1 + 1}
    for k, v in ms.items():
        v /= (Fs + (1 << 18) - 1) // (1 << 18)
        print(f"staged MC {k:8s} {v:8.4f} ms per 2^18 frames  {algb[k] * (1 << 18) / v / 1e6:8.1f} GB/s")
