"""Oracle: whole Monte-Carlo step on the CPU (TEST INFRASTRUCTURE / CPU baseline).

One step = what /root/reference/src/fnn_viterbi.py does between `train_test_dataset`
(test half :380-388), the test-set forward (:606-610) and `calculate_error` (:619-674):
  class-balanced labels -> encode_133171 -> BPSK+AWGN -> MLP forward -> argmax -> count.

Two noise sources:
  * "philox": bit-compatible with the CUDA path (checker for counters / decisions);
  * "numpy" : vectorised numpy Gaussian draws — the fastest honest CPU port of the
    reference's `np.random.normal` call, used only when the oracle is TIMED as the
    CPU baseline (bench.py cpu_baseline / --impl reference).
The MLP runs in torch-CPU fp32 (MKL, all host threads) when torch is importable — the
stand-in for the reference's TensorFlow CPU session — else in numpy fp32.
"""
from __future__ import annotations

import numpy as np

from . import channel, encoders, metrics, mlp, philox


def frame_labels(frame0: int, B: int, L: int) -> np.ndarray:
    """Class label of global frame f is f mod 2^L (class-balanced like fnn_viterbi.py:380-388)."""
    return (np.arange(frame0, frame0 + B, dtype=np.int64) & ((1 << L) - 1))


def codebook(L: int, code: int) -> np.ndarray:
    """full_encoded_set of the reference (fnn_viterbi.py:347-352): [2^L, 2L] in {0,1}."""
    return encoders.conv_encode(encoders.generate_class_bits(L), code)


def _forward_fast(y32, Ws, bs, acts, threads=None):
    try:
        import torch
    except Exception:  # pragma: no cover
        return mlp.forward(y32, Ws, bs, acts, dtype=np.float32)
    if threads:
        torch.set_num_threads(threads)
    with torch.no_grad():
        h = torch.from_numpy(np.ascontiguousarray(y32, dtype=np.float32))
        for W, b, a in zip(Ws, bs, acts):
            h = torch.addmm(torch.from_numpy(np.asarray(b, np.float32)), h,
                            torch.from_numpy(np.asarray(W, np.float32)))
            if a == mlp.ACT_TANH:
                h = torch.tanh_(h)
            elif a == mlp.ACT_RELU:
                h = torch.relu_(h)
        return h.numpy()


def mc_step(Ws, bs, acts, L, code, snr_db, frame0, B, seed=philox.DEFAULT_SEED, snr_idx=0,
            noise="philox", y_dtype=channel.OUT_F32, precision="fp32", rng=None,
            return_detail=False):
    """Returns counters [bit_errors, frame_errors, frames, bits] (and details)."""
    sigma = float(channel.snr_db_to_sigma(snr_db))
    labels = frame_labels(frame0, B, L)
    coded = codebook(L, code)[labels]
    if noise == "philox":
        y = channel.channel_philox(coded, sigma, seed, snr_idx, frame0, y_dtype)
    else:
        gen = rng if rng is not None else np.random.default_rng(seed)
        z = gen.standard_normal(coded.shape, dtype=np.float32)
        y = (2.0 * coded.astype(np.float32) - 1.0) + np.float32(sigma) * z
    if precision == "fp64":
        out = mlp.forward(y, Ws, bs, acts, dtype=np.float64)
    elif precision == "bf16":
        out, _ = mlp.forward_bf16_model(y, Ws, bs, acts)
    else:
        out = _forward_fast(y.astype(np.float32), Ws, bs, acts)
    pred = mlp.argmax_first(out)
    be, fe = metrics.count_errors(pred, labels, L)
    counters = np.array([be, fe, B, B * L], dtype=np.uint64)
    if return_detail:
        return counters, dict(y=y, out=out, pred=pred, labels=labels)
    return counters
