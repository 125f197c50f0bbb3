#!/usr/bin/env python
"""Multi-GPU check of the peer-memory exchange (csrc/comm.cu): run under torchrun on >= 2 GPUs of one node.
Every rank contributes rank- and round-dependent values; the sums must be exact on every rank, round after round
(the two alternating areas), for both transports; then both are timed on the device."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "neural-network-based-viterbi-decoding_b200"))
from nvd import dist, montecarlo

rank, world, local = dist.init()
res = {}
for transport in ("p2p", "nccl"):
    os.environ["NVD_P2P"] = "1" if transport == "p2p" else "0"
    with montecarlo.Simulator(device=local) as sim:
        sim.comm_init(rank, world)
        info = sim.comm_info()
        assert (info["transport"] == "nvlink-peer-memory") == (transport == "p2p"), info
        ok = True
        for rnd in range(12):
            n = [4, 36, 448, 1][rnd % 4]
            mine = (np.arange(n, dtype=np.uint64) * np.uint64(1000003) + np.uint64(rank * 7919 + rnd * 104729 + 1))
            buf = sim.to_device(mine)
            sim.allreduce_counters(buf, n)
            got = buf.download(np.uint64, (n,))
            want = sum((np.arange(n, dtype=np.uint64) * np.uint64(1000003) + np.uint64(r * 7919 + rnd * 104729 + 1)) for r in range(world))
            ok = ok and np.array_equal(got, want)
            sim.free(buf)
        buf = sim.zeros(32)
        for _ in range(20):
            sim.allreduce_counters(buf, 4)
        sim.sync(); dist.barrier()
        e0, e1 = sim.event(), sim.event()
        sim.record(e0)
        for _ in range(200):
            sim.allreduce_counters(buf, 4)
        sim.record(e1); sim.sync()
        us = sim.elapsed_ms(e0, e1) / 200 * 1e3
        res[transport] = (ok, dist.all_reduce_max(us))
        dist.barrier()
if rank == 0:
    for t, (ok, us) in res.items():
        print(f"{t:5s}: sums {'exact' if ok else 'WRONG'} on rank 0; {us:.1f} us per all-reduce of 4 x uint64 over {world} GPUs (max over ranks)")
oks = dist.all_reduce_max(0.0 if all(v[0] for v in res.values()) else 1.0)
dist.shutdown()
sys.exit(0 if oks == 0.0 else 1)
