#!/usr/bin/env python
"""Staged kernels K0 / K1 / K2 / K5 at growing batch sizes: algorithmic bytes / CUDA-event time against
the measured HBM copy bandwidth (MEASURED_PEAKS.json).  At 2^24 frames the buffers are 17-470 MB and the kernels last
10-30 us (launch-latency sized); the larger sizes show the asymptote."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "neural-network-based-viterbi-decoding_b200"))
from nvd import montecarlo, _lib

peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
L, n = 7, 14
with montecarlo.Simulator(0) as sim:
    for logB in (24, 26, 28):
        B = 1 << logB
        info = sim.alloc(B); coded = sim.alloc(2 * B); y = sim.alloc(B * n * 2); pred = sim.alloc(B); cnt = sim.zeros(32)
        steps = {
            "K0 source": (lambda: _lib.check(sim.lib.nvd_source(sim.ctx, L, 0, B, info.ptr)), 1),
            "K1 encode": (lambda: _lib.check(sim.lib.nvd_encode(sim.ctx, _lib.CODE_133171, info.ptr, L, B, coded.ptr)), 3),
            "K2 channel bf16": (lambda: _lib.check(sim.lib.nvd_channel(sim.ctx, coded.ptr, n, B, 0.5, 0, 0, _lib.BF16, y.ptr)), 30),
            "K5 count": (lambda: _lib.check(sim.lib.nvd_count_errors(sim.ctx, pred.ptr, info.ptr, 1, L, B, 0, cnt.ptr)), 2),
        }
        for name, (fn, bytes_per_frame) in steps.items():
            fn(); sim.sync()
            e0, e1 = sim.event(), sim.event()
            sim.record(e0)
            for _ in range(5):
                fn()
            sim.record(e1); sim.sync()
            ms = sim.elapsed_ms(e0, e1) / 5
            gbs = bytes_per_frame * B / (ms * 1e-3) / 1e9
            print(f"B=2^{logB}  {name:16s} {ms * 1e3:9.1f} us  {gbs:8.1f} GB/s  {gbs / peak:6.1%} of measured HBM copy bandwidth")
        for b in (info, coded, y, pred):
            sim.free(b)
