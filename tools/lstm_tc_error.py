import os, sys
import numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/neural-network-based-viterbi-decoding_b200")
from nvd import montecarlo, weights
from oracle import lstm as olstm, encoders as oenc
H, L, B = 20, 100, 1000
p = olstm.init_params(H, seed=4)
rng = np.random.default_rng(9)
p["enc_b"] = (0.1 * rng.standard_normal(4 * H)).astype(np.float32); p["dec_b"] = (0.1 * rng.standard_normal(4 * H)).astype(np.float32)
info = rng.integers(0, 2, (B, L))
y = (2.0 * oenc.encode_133171(info) - 1.0 + 0.6 * rng.standard_normal((B, 2 * L))).astype(np.float32)
_, dec_in, _ = olstm.prepare_batch(2.0 * info - 1.0, y, L)
want = olstm.seq2seq_forward(y, dec_in, p, dtype=np.float64)
with montecarlo.Simulator(0) as sim:
    h = sim.load_lstm(weights.LstmModel(**p))
    lb, _ = sim.lstm_forward(h, sim.to_device(y), "f32", sim.to_device(dec_in.astype(np.float32)), 2 * L, 2 * L, B, fast_math=2)
    got = lb.download(np.float32, (B, 2 * L))
    e = np.abs(got - want)
    print("max err %.4f  rms %.5f  max|logit| %.3f" % (e.max(), np.sqrt((e ** 2).mean()), np.abs(want).max()))
