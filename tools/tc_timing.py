#!/usr/bin/env python
"""Per-phase clock64 timeline of the fused tcgen05 kernel (CTA 0, its groups, first 64 tiles) and the finish times of
all groups of the launch (how evenly the static deal of tiles ends)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "neural-network-based-viterbi-decoding_b200"))
from nvd import montecarlo, weights, _lib

G = int(os.environ.get("NVD_TC_GROUPS", "2"))      # groups per CTA of the build
names = ["top", "-", "-", "wait1", "epi1", "handoff2", "gen1 (Philox)", "wait2",
         "epi2+handoff3", "gen2 (radii)", "wait3", "epi3+handoff4", "gen3 (symbols)", "wait4", "argmax"]
STAMPS = [0, 1, 2, 3, 4, 5, 8, 9, 10, 6, 11, 12, 7, 13, 14]
with montecarlo.Simulator(0) as sim:
    h = sim.load_mlp(weights.load_default_fnn())
    raw = np.zeros(8192, dtype=np.uint64)
    c = sim.zeros(32)
    for s_ in range(30):                                     # warm clocks: the stamped launch follows ~28 ms of full load
        sim.run_async(h, "133171", 7, [2.0], 1 << 24, c, s_ << 24, mode="fused")
    _lib.check(sim.lib.nvd_debug_tc_timing(sim.ctx, h, 1 << 24, raw.ctypes.data))
    st = raw[:64 * G * 16].reshape(64, G, 16).astype(np.int64)
    d = np.diff(st[:, :, STAMPS], axis=2)                    # [tile, group, phases] durations
    per_tile = st[1:, :, 0] - st[:-1, :, 0]
    print("tile period per group (clk): median", np.median(per_tile[8:], axis=0), " mean", per_tile[8:].mean())
    med = np.median(d[8:56], axis=(0, 1))
    for n, v in zip(names[1:], med):
        print(f"  {n:22s} {v:8.0f}")
    print("  sum", med.sum())
    fin = raw[4096:].reshape(-1, 4).astype(np.int64)
    fin = fin[fin[:, 2] > 0]
    t0 = fin[:, 0].min()
    start, end, tiles, smid = fin[:, 0] - t0, fin[:, 1] - t0, fin[:, 2], fin[:, 3]
    print(f"groups {len(fin)}: tiles min/max {tiles.min()}/{tiles.max()}; start spread {start.max()/1e3:.1f} us; "
          f"finish min {end.min()/1e3:.1f} us, median {np.median(end)/1e3:.1f} us, max {end.max()/1e3:.1f} us")
    per = (end - start) / tiles
    sms, counts = np.unique(smid, return_counts=True)
    print(f"SMs used {len(sms)}; groups per SM: " + ", ".join(f"{c} groups on {n} SMs" for c, n in zip(*np.unique(counts, return_counts=True))))
    for c in np.unique(counts):
        sel = np.isin(smid, sms[counts == c])
        print(f"  SMs with {c} groups: ns per tile per group mean {per[sel].mean():.1f} (min {per[sel].min():.1f}, max {per[sel].max():.1f}); "
              f"finish mean {end[sel].mean()/1e3:.1f} us; SM tile rate {c / per[sel].mean() * 1e3:.3f} tiles/us")
    idle = (end.max() - end).sum() / (end.max() * len(end))
    print(f"capacity idle behind finished groups: {100*idle:.2f} % of the launch")
