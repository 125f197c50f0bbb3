"""Oracle: BPSK + AWGN channel (TEST INFRASTRUCTURE).

Restates modulate_awgn (/root/reference/src/fnn_viterbi.py:224-253; four more copies in
fnn_viterbi_mixed_snr.py, fnn_viterbi_bitwise.py:223-237, fnn_viterbi_seq_predict.py,
rnn_viterbi.py:151-165): x = float32(c) with 0 -> -1, then y = x + N(0, sigma^2).
The destination dtype is the caller's: float32 rows (fnn_viterbi.py:373,384) or int32
rows, i.e. truncation toward zero (fnn_viterbi_bitwise.py:309-317, rnn_viterbi.py:201-208).
"""
from __future__ import annotations

import numpy as np

from . import philox

OUT_F32 = 0
OUT_BF16 = 1
OUT_I32_TRUNC = 2


def snr_db_to_sigma(snr_db) -> np.ndarray:
    """sigma = sqrt(1 / (2 * 10^(dB/10)))  (fnn_viterbi.py:89-92; Es/N0 per coded symbol)."""
    snr = 10.0 ** (np.asarray(snr_db, dtype=np.float64) / 10.0)
    return np.sqrt((1.0 / snr) * 0.5)


def bpsk(coded) -> np.ndarray:
    """0 -> -1, 1 -> +1 as float32 (fnn_viterbi.py:244-247)."""
    x = np.asarray(coded).astype(np.float32)
    x = np.where(x == 0, np.float32(-1.0), x)
    return x


def modulate_awgn(encoded_seq, sigma: float, rng=None) -> np.ndarray:
    """Reference-semantics channel: noise from numpy's legacy global RNG unless `rng`
    (anything with .normal) is given.  Returns float64 like the reference (:250-251)."""
    x = bpsk(encoded_seq)
    gen = np.random if rng is None else rng
    return x + gen.normal(0, sigma, x.shape[-1] if x.ndim == 1 else x.shape)


def bf16_round(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even fp32 -> bf16 -> fp32 (what cvt.rn.bf16.f32 does)."""
    u = np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + np.uint64(0x7FFF) + ((u >> np.uint64(16)) & np.uint64(1))) >> np.uint64(16)) << np.uint64(16)
    return r.astype(np.uint32).view(np.float32)


def channel_philox(coded: np.ndarray, sigma: float, seed: int, snr_idx: int,
                   frame0: int = 0, out_dtype: int = OUT_F32) -> np.ndarray:
    """Deterministic channel used by the B200 path: y = (2c-1) + sigma * z with z from
    Philox (oracle/philox.py).  coded: [B, n] in {0,1}; frame b has global index frame0+b."""
    coded = np.asarray(coded)
    B, n = coded.shape
    frames = np.arange(frame0, frame0 + B, dtype=np.uint64)
    z = philox.noise_normals(seed, snr_idx, frames, n)
    y = (2.0 * coded.astype(np.float64) - 1.0) + np.float64(np.float32(sigma)) * z
    if out_dtype == OUT_F32:
        return y.astype(np.float32)
    if out_dtype == OUT_BF16:
        return bf16_round(y.astype(np.float32))
    if out_dtype == OUT_I32_TRUNC:
        return np.trunc(y).astype(np.int32)
    raise ValueError(out_dtype)
