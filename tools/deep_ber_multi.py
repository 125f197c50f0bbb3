#!/usr/bin/env python
"""BASELINE.json configs[4]: deep-BER Monte-Carlo run of the (133,171) FNN decoder, frames sharded over the ranks of
one box (torchrun), counters combined by one NCCL all-reduce per chunk.  Stop rule per SNR point: >= E frame errors
(global) or N_max frames.  Also checks shard invariance at scale: the N-rank counters of a fixed frame range equal the
single-rank counters of the same range.  Writes gpurun_out/deep_ber_n<world>.json (rank 0).
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/deep_ber_multi.py"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "neural-network-based-viterbi-decoding_b200"))
from nvd import dist, montecarlo, weights

rank, world, local = dist.init()
m = weights.load_default_fnn()
snrs = [6.0, 7.0, 8.0, 9.0]
E, N_MAX, CHUNK = 2000, 1 << 38, 1 << 32
out = {"world": world, "snr_db": snrs, "min_frame_errors": E, "chunk_frames": CHUNK, "points": []}
with montecarlo.Simulator(local) as sim:
    h = sim.load_mlp(m)
    ar = dist.all_reduce_sum if world > 1 else None
    # shard invariance on real NCCL: a fixed range, all ranks vs rank 0 alone
    fixed = sim.run(h, "133171", 7, [4.0, 6.0], 1 << 28, 12345, "fused", "bf16", None, None, rank, world, ar)
    if rank == 0:
        alone = sim.run(h, "133171", 7, [4.0, 6.0], 1 << 28, 12345, "fused")
        out["shard_invariance"] = {"frames": 1 << 28, "sharded": fixed.tolist(), "single_rank": alone.tolist(),
                                   "equal": bool((fixed == alone).all())}
    dist.barrier()
    t0 = time.perf_counter()
    total = sim.sweep(h, "133171", 7, snrs, min_frame_errors=E, max_frames=N_MAX, chunk_frames=CHUNK, mode="fused",
                      rank=rank, world=world, all_reduce=ar)
    dt = time.perf_counter() - t0
    if rank == 0:
        frames = int(total[:, 2].sum())
        for db, c in zip(snrs, total):
            lo, hi = montecarlo.wilson(float(c[0]), float(c[3]))
            out["points"].append({"snr_db": db, "frames": int(c[2]), "bit_errors": int(c[0]), "frame_errors": int(c[1]),
                                  "ber": float(c[0]) / float(c[3]), "ber_ci95": [lo, hi]})
        out["seconds"] = dt
        out["frames_total"] = frames
        out["gbit_per_s_incl_host"] = frames * 7 / dt / 1e9
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"deep_ber_n{world}.json"), "w"), indent=1)
        print(json.dumps(out))
dist.shutdown()
