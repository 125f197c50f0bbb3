import os, sys
import numpy as np
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "neural-network-based-viterbi-decoding_b200"))
os.environ["NVD_LIB"] = os.path.join(ROOT, "neural-network-based-viterbi-decoding_b200/build/var/libnvd_DEBUG_WARP_STAMPS=1.so")
from nvd import montecarlo, weights, _lib
with montecarlo.Simulator(0) as sim:
    h = sim.load_mlp(weights.load_default_fnn())
    st = np.zeros(64 * 4 * 16, dtype=np.uint64)
    _lib.check(sim.lib.nvd_debug_tc_timing(sim.ctx, h, 1 << 24, st.ctypes.data))
    st = st.reshape(64, 4, 16).astype(np.int64)[8:56]
    rel = st[:, :, 10:14] - st[:, :, 3:4]
    print("mid arrive of warps 0..3 after epi1 start (median):", np.median(rel, axis=(0, 1)).tolist(), " max-min spread median:", np.median(rel.max(2) - rel.min(2)))
    print("rdy0 seen by MMA warp after the LAST warp's mid arrive (median):", np.median(st[:, :, 6] - st[:, :, 10:14].max(2)))
    print("epi1 end (warp 0) after epi1 start:", np.median(st[:, :, 4] - st[:, :, 3]))
