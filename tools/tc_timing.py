#!/usr/bin/env python
"""Per-phase clock64 timeline of the fused tcgen05 kernel (CTA 0, its 4 groups, first 64 tiles)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "neural-network-based-viterbi-decoding_b200"))
from nvd import montecarlo, weights, _lib

names = ["top", "-", "-", "wait1", "epi1", "handoff2", "gen1 (Philox)", "wait2",
         "epi2+handoff3", "gen2 (radii)", "wait3", "epi3+handoff4", "gen3 (symbols)", "wait4", "argmax"]
STAMPS = [0, 1, 2, 3, 4, 5, 8, 9, 10, 6, 11, 12, 7, 13, 14]
with montecarlo.Simulator(0) as sim:
    h = sim.load_mlp(weights.load_default_fnn())
    st = np.zeros(64 * 4 * 16, dtype=np.uint64)
    _lib.check(sim.lib.nvd_debug_tc_timing(sim.ctx, h, 1 << 24, st.ctypes.data))
    st = st.reshape(64, 4, 16).astype(np.int64)
    t0 = st[:, :, 0].min()
    d = np.diff(st[:, :, STAMPS], axis=2)                    # [tile, group, 14] phase durations
    per_tile = st[1:, :, 0] - st[:-1, :, 0]
    print("tile period per group (clk): median", np.median(per_tile[8:], axis=0), " mean", per_tile[8:].mean())
    med = np.median(d[8:56], axis=(0, 1))
    for n, v in zip(names[1:], med):
        print(f"  {n:22s} {v:8.0f}")
    print("  sum", med.sum())
    print("group phase offsets at tile 20 (top stamp rel. to group 0):", (st[20, :, 0] - st[20, 0, 0]).tolist())
    print("group phase offsets at tile 40:", (st[40, :, 0] - st[40, 0, 0]).tolist())
