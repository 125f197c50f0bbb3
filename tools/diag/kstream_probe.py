#!/usr/bin/env python
"""Diagnostic: expose the K-streamed kernel's first-layer accumulators (identity activations, W2 = I) and compare them
column by column with exact bf16 x bf16 products summed in float64."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "neural-network-based-viterbi-decoding_b200"))
from nvd import montecarlo as mc, weights
from oracle import channel as ochan

rng = np.random.default_rng(0)
with mc.Simulator(0) as sim:
    for K, N1, real in ((48, 8, True), (48, 8, False), (96, 8, True), (1036, 8, True), (48, 30, True)):
        n_out = 8
        W1 = (rng.standard_normal((K, N1)) * np.sqrt(2.0 / K)).astype(np.float32)
        b1 = (0.05 * rng.standard_normal(N1)).astype(np.float32)
        W2 = np.zeros((N1, n_out), np.float32); W2[np.arange(8), np.arange(8)] = 1.0
        b2 = np.zeros(n_out, np.float32)
        h = sim.load_mlp(weights.MlpModel([W1, W2], [b1, b2], [0, 0]))
        B = 256
        y = rng.standard_normal((B, K)).astype(np.float32) if real else np.trunc(2 * rng.standard_normal((B, K))).astype(np.float32)
        y16 = mc.f32_to_bf16(y); yq = mc.bf16_to_f32(y16).astype(np.float64)
        soft_b, _ = sim.mlp_forward(h, sim.to_device(y16), "bf16", B, n_out, "bf16", want_soft=True, hard_bytes=4)
        soft = soft_b.download(np.float32, (B, n_out))
        ref = yq @ ochan.bf16_round(W1).astype(np.float64)[:, :8] + b1[:8]
        err = np.abs(soft - ref)
        print(f"K={K} N1={N1} real={real}: max err per column {np.round(err.max(0), 6).tolist()}  rows with err>1e-4: {(err.max(1) > 1e-4).sum()}")
        # which k contributes? single-hot rows
        if K == 48 and N1 == 8 and real:
            e = np.eye(K, dtype=np.float32)
            sb, _ = sim.mlp_forward(h, sim.to_device(mc.f32_to_bf16(e)), "bf16", K, n_out, "bf16", want_soft=True, hard_bytes=4)
            got = sb.download(np.float32, (K, n_out)) - b1[:8]
            d = np.abs(got - ochan.bf16_round(W1)[:, :8])
            print("  unit rows: max err", d.max(), "bad k:", np.nonzero(d.max(1) > 1e-6)[0].tolist())
