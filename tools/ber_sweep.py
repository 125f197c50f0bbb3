#!/usr/bin/env python
"""BER/FER sweep of the (133,171) FNN decoder: fused tcgen05 pipeline vs the CPU oracle on the SAME
Philox frames (parity), and a deep run (stop rule: >= 100 frame errors or max frames) per SNR point.
Writes gpurun_out/ber_sweep.json."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "neural-network-based-viterbi-decoding_b200"))
from nvd import montecarlo, weights
from oracle import pipeline as opipe, encoders as oenc, metrics as omet

m = weights.load_default_fnn()
snrs = list(range(0, 9))
out = {"weights": m.meta.get("path"), "seed": montecarlo.DEFAULT_SEED, "points": []}
ML_BER = [1.77e-2, 7.09e-3, 2.37e-3, 6.72e-4, 1.66e-4, 3.3e-5, 5.5e-6, 6.4e-7, 3.6e-8]
with montecarlo.Simulator(0) as sim:
    h = sim.load_mlp(m)
    n_par = 1 << 18
    t0 = time.time()
    gpu_par = sim.run(h, "133171", 7, snrs, n_par, 0, mode="fused")
    fp32_par = sim.run(h, "133171", 7, snrs, n_par, 0, mode="staged", precision="fp32", y_dtype="f32")
    deep = sim.sweep(h, "133171", 7, snrs, min_frame_errors=100, max_frames=1 << 34, chunk_frames=1 << 28, mode="fused")
    t_gpu = time.time() - t0
    for i, db in enumerate(snrs):
        ref = opipe.mc_step(m.Ws, m.bs, m.acts, 7, oenc.CODE_133171, float(db), 0, n_par, snr_idx=i, noise="philox", precision="fp32")
        lo, hi = omet.wilson_interval(int(ref[0]), int(ref[3]))
        d = deep[i]
        dlo, dhi = montecarlo.wilson(float(d[0]), float(d[3]))
        out["points"].append({
            "snr_db": db, "frames_parity": n_par,
            "oracle_fp32": [int(x) for x in ref], "gpu_fp32": [int(x) for x in fp32_par[i]], "gpu_bf16_fused": [int(x) for x in gpu_par[i]],
            "oracle_ber": float(ref[0]) / float(ref[3]), "oracle_ber_ci95": [lo, hi],
            "gpu_ber_same_frames": float(gpu_par[i][0]) / float(gpu_par[i][3]),
            "gpu_inside_oracle_ci": bool(lo <= float(gpu_par[i][0]) / float(gpu_par[i][3]) <= hi),
            "deep": {"frames": int(d[2]), "bit_errors": int(d[0]), "frame_errors": int(d[1]), "ber": float(d[0]) / float(d[3]),
                     "fer": float(d[1]) / float(d[2]), "ber_ci95": [dlo, dhi]},
            "ml_bound_ber": ML_BER[i],
        })
        print(out["points"][-1], flush=True)
out["gpu_seconds"] = t_gpu
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "ber_sweep.json"), "w"), indent=1)
