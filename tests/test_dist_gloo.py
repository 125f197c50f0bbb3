"""world_size-2 gloo test of the N>1 host path: frame-range sharding + the counter all-reduce.
The per-rank 'device work' is played by the CPU oracle here (no GPU in this container); on the GPU
box tests/test_gpu_pipeline.py::test_shard_invariance checks the same property with the kernels."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys, json
    import numpy as np
    sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "neural-network-based-viterbi-decoding_b200"))
    from nvd import dist, montecarlo as mc
    from oracle import pipeline as opipe, mlp as omlp, encoders as oenc
    rank, world, local = dist.init("gloo")
    Ws, bs = omlp.init_fnn([14, 128, 64, 32, 128], seed=3, bias_scale=0.05)
    N, frame0 = 6001, 17
    lo, n = mc.shard_range(frame0, N, rank, world)
    c = opipe.mc_step(Ws, bs, [1, 1, 1, 1], 7, oenc.CODE_133171, 2.0, lo, n, noise="philox")
    tot = dist.all_reduce_sum(c.astype(np.int64))
    worst = dist.all_reduce_max(float(rank))
    # the byte channels of the exchange set-up: NCCL id from rank 0, CUDA IPC handles from everybody in rank order
    ident = dist.broadcast_bytes(bytes(range(128)) if rank == 0 else None, 128)
    handles = dist.all_gather_bytes(bytes([rank * 16 + i for i in range(64)]), 64)
    dist.barrier()
    if rank == 0:
        whole = opipe.mc_step(Ws, bs, [1, 1, 1, 1], 7, oenc.CODE_133171, 2.0, frame0, N, noise="philox")
        print(json.dumps(dict(tot=tot.tolist(), whole=whole.astype(np.int64).tolist(), worst=worst,
                              ident_ok=ident == bytes(range(128)),
                              handles_ok=handles == bytes([r * 16 + i for r in range(world) for i in range(64)]))))
    dist.shutdown()
""")


def test_two_rank_sharding_and_allreduce(tmp_path):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), OMP_NUM_THREADS="2")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=240) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    import json
    res = json.loads(outs[0][0].strip().splitlines()[-1])
    assert res["tot"] == res["whole"]          # counters independent of the split
    assert res["worst"] == 1.0
    assert res["ident_ok"] and res["handles_ok"]
    assert res["tot"][2] == 6001 and res["tot"][3] == 6001 * 7
