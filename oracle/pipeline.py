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


def long_frames(L, frame0, B, seed=philox.DEFAULT_SEED):
    """Philox info bits [B, L] of global frames frame0 .. frame0+B-1 (what K0 writes for L > 16;
    generate_input_vectors, /root/reference/src/fnn_viterbi_bitwise.py:239-250, draws them from numpy's RNG)."""
    return philox.info_bits(seed, np.arange(frame0, frame0 + B, dtype=np.uint64), L)


def mc_step_bitwise(Ws, bs, acts, L, snr_db, frame0, B, seed=philox.DEFAULT_SEED, snr_idx=0, state=0,
                    n_state_bits=6, bit_position=0, trunc=True, precision="fp64", code=encoders.CODE_133171,
                    return_detail=False):
    """The per-bit predictor's test (fnn_viterbi_bitwise.py:290-323 data, :388-397 net, :456-457 label, :504-509
    decision and error rate) on Philox frames: counters [errors, errors, frames, frames]."""
    info = long_frames(L, frame0, B, seed)
    prefix = encoders.extract_state_bits(state, n_state_bits) if n_state_bits else np.zeros(0, np.uint8)
    full = np.concatenate([np.tile(prefix.astype(np.uint8), (B, 1)), info], axis=1)
    coded = encoders.conv_encode(full, code)
    sigma = float(channel.snr_db_to_sigma(snr_db))
    dt = channel.OUT_I32_TRUNC if trunc else (channel.OUT_BF16 if precision == "bf16" else channel.OUT_F32)
    y = channel.channel_philox(coded, sigma, seed, snr_idx, frame0, dt).astype(np.float32)
    if precision == "bf16":
        Wq = [channel.bf16_round(np.asarray(Ws[0], np.float32))] + [np.asarray(W, np.float32) for W in Ws[1:]]
        out = mlp.forward(channel.bf16_round(y), Wq, bs, acts, dtype=np.float64)     # bf16 inputs / W1, fp32 tail
    else:
        out = mlp.forward(y, Ws, bs, acts, dtype=np.float64 if precision == "fp64" else np.float32)
    pred = mlp.argmax_first(out)
    err = int((pred != info[:, bit_position]).sum())
    counters = np.array([err, err, B, B], dtype=np.uint64)
    if return_detail:
        return counters, dict(y=y, out=out, pred=pred, info=info, coded=coded)
    return counters


def mc_step_lstm(params, L, snr_db, frame0, B, seed=philox.DEFAULT_SEED, snr_idx=0, teacher_forced=True, trunc=True,
                 code=encoders.CODE_133171, return_detail=False):
    """The recurrent decoder's test (rnn_viterbi.py:186-232 data, :328-377 batch, :278-305 forward, :471-491 BER over
    the first L decoder steps) on Philox frames, decision bit = logit > 0 (Appendix B-12)."""
    from . import lstm
    info = long_frames(L, frame0, B, seed)
    coded = encoders.conv_encode(info, code)
    sigma = float(channel.snr_db_to_sigma(snr_db))
    y = channel.channel_philox(coded, sigma, seed, snr_idx, frame0,
                               channel.OUT_I32_TRUNC if trunc else channel.OUT_F32).astype(np.float64)
    _, dec_in, _ = lstm.prepare_batch(2.0 * info - 1.0, y, L)
    logits = lstm.seq2seq_forward(y, dec_in, params, dtype=np.float64, autoregressive=not teacher_forced)
    pred = (logits[:, :L] > 0).astype(np.uint8)
    x = pred ^ info
    counters = np.array([int(x.sum()), int(x.any(axis=1).sum()), B, B * L], dtype=np.uint64)
    if return_detail:
        return counters, dict(y=y, logits=logits, pred=pred, info=info)
    return counters
