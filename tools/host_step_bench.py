#!/usr/bin/env python
"""Host-buffer step (nvd_mc_run_host, fused kernel, pinned buffers): sliced launches vs the single streaming launch,
with decisions and counters only; checks that both forms return the same decisions and counters."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "neural-network-based-viterbi-decoding_b200"))
from nvd import montecarlo, weights

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
reps = 20
with montecarlo.Simulator(0) as sim:
    h = sim.load_mlp(weights.load_default_fnn())
    info, pred = sim.pinned(frames), sim.pinned(frames)
    info[:] = np.random.default_rng(1).integers(0, 128, frames).astype(np.uint8)
    ref = {}
    for mode in ("sliced", "streaming", "sliced", "streaming"):
        os.environ["NVD_HOST_STEP"] = mode
        for want_pred in (True, False):
            pred[:] = 255
            c = sim.run_host(h, "133171", 7, 2.0, info, pred if want_pred else None, frame0=0, mode="fused")
            t0 = time.perf_counter()
            for _ in range(reps):
                c = sim.run_host(h, "133171", 7, 2.0, info, pred if want_pred else None, frame0=0, mode="fused")
            ms = (time.perf_counter() - t0) / reps * 1e3
            key = "pred" if want_pred else "counters"
            ok = ""
            if key in ref:
                ok = "  == first" if (ref[key][0] == c).all() and (not want_pred or (ref[key][1] == pred).all()) else "  MISMATCH"
            else:
                ref[key] = (c.copy(), pred.copy() if want_pred else None)
            print(f"{mode:9s} {key:8s}: {ms:.4f} ms/step  {frames * 7 / ms / 1e6:.2f} Gbit/s  counters {c.tolist()}{ok}")
    os.environ.pop("NVD_HOST_STEP")
