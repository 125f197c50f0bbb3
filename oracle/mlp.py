"""Oracle: feed-forward decoder graphs (TEST INFRASTRUCTURE, parity UNPINNED by the reference).

Restates the graph DEFINITIONS; the arithmetic itself ran inside TensorFlow 1.x
(`tensorflow>=2.0.0,<3.0.0` pinned in requirements.txt:2, code needs <=1.15), which is not
under /root/reference and not installable here:
  fnn_viterbi / mixed_snr : out = tanh(tanh(tanh(tanh(xW1+b1)W2+b2)W3+b3)Wout+bout)
                            /root/reference/src/fnn_viterbi.py:495-508, fnn_viterbi_mixed_snr.py:589-592
  bitwise                 : out = relu(relu(xW1+b1)Wout+bout)   fnn_viterbi_bitwise.py:388,397
  seq_predict             : tanh after all five affines           fnn_viterbi_seq_predict.py:408-412
  decision                : tf.argmax(out, 1), first index on ties  fnn_viterbi.py:606-610
  init                    : truncated_normal(stddev=sqrt(2/n_in)), zero biases  fnn_viterbi.py:430-466
Weights are [in, out], applied as x @ W + b.
"""
from __future__ import annotations

import numpy as np

ACT_IDENTITY = 0
ACT_TANH = 1
ACT_RELU = 2


def _act(z, a):
    if a == ACT_TANH:
        return np.tanh(z)
    if a == ACT_RELU:
        return np.maximum(z, 0)
    return z


def forward(x, Ws, bs, acts, dtype=np.float64, return_preact_last=False):
    """Generic MLP forward in `dtype`; returns the final activations [B, n_out]."""
    h = np.asarray(x, dtype=dtype)
    z = h
    for W, b, a in zip(Ws, bs, acts):
        z = h @ np.asarray(W, dtype=dtype) + np.asarray(b, dtype=dtype)
        h = _act(z, a)
    return (h, z) if return_preact_last else h


def tanh_two_pipe(z, tanh_model):
    """tanh as the tensor-core kernel evaluates it for hidden layers (include/nvd.h,
    nvd_tc_tanh_model): of every 8 column pairs the first `n` use the fp32 polynomial
    sign(z) u Q(u), u = min(|z| inv_range, 1); the others exact tanh (tanh.approx)."""
    n, coef, inv_range = tanh_model
    z = np.asarray(z, dtype=np.float32)
    out = np.tanh(z.astype(np.float64)).astype(np.float32)
    cols = np.arange(z.shape[1])
    poly_cols = ((cols // 2) % 8) < n
    u = np.minimum(np.abs(z[:, poly_cols]) * np.float32(inv_range), np.float32(1.0)).astype(np.float32)
    p = np.full_like(u, np.float32(coef[5]))
    for c in coef[4::-1]:
        p = (p.astype(np.float64) * u + np.float32(c)).astype(np.float32)     # one rounding per fma
    p = (p * u).astype(np.float32)
    out[:, poly_cols] = np.copysign(p, z[:, poly_cols])
    return out


def forward_bf16_model(x, Ws, bs, acts, tanh_model=None):
    """Numerical model of the tensor-core path: inputs, weights and inter-layer activations
    rounded to bf16, fp32 accumulation, biases bf16 for hidden layers (they are folded
    into the contraction); hidden tanh exact, or the kernel's two-pipe evaluation when
    `tanh_model` = (poly_pairs_of_8, coef, inv_range) is given.  Used to bound the CUDA bf16
    kernel tighter than the 2e-2 north-star tolerance."""
    from .channel import bf16_round
    h = bf16_round(np.asarray(x, dtype=np.float32))
    z = None
    for li, (W, b, a) in enumerate(zip(Ws, bs, acts)):
        Wq = bf16_round(np.asarray(W, dtype=np.float32)).astype(np.float64)
        bq = bf16_round(np.asarray(b, dtype=np.float32)).astype(np.float64)
        z = (h.astype(np.float64) @ Wq + bq).astype(np.float32)
        if tanh_model is not None and a == ACT_TANH and li != len(Ws) - 1:
            h = tanh_two_pipe(z, tanh_model)
        else:
            h = _act(z, a).astype(np.float32)
        if li != len(Ws) - 1:
            h = bf16_round(h)
    return h, z


def argmax_first(out) -> np.ndarray:
    """Class of the largest output, first index on ties (numpy and TF agree)."""
    return np.argmax(out, axis=1).astype(np.int64)


def top2_margin(out) -> np.ndarray:
    """top-1 minus top-2 output per row — the 'logit margin' of DESIGN.md."""
    part = np.partition(np.asarray(out, dtype=np.float64), -2, axis=1)
    return part[:, -1] - part[:, -2]


def truncated_normal(shape, stddev, rng: np.random.Generator) -> np.ndarray:
    """TF-1.x tf.truncated_normal: redraw anything beyond 2 stddev (fnn_viterbi.py:432)."""
    out = rng.standard_normal(shape)
    bad = np.abs(out) > 2.0
    while bad.any():
        out[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(out) > 2.0
    return (out * stddev).astype(np.float32)


def init_fnn(dims, seed: int = 0, bias_scale: float = 0.0):
    """Seed-fixed weights with the reference's init distribution; `bias_scale` > 0 fills the
    biases with small non-zero values so parity tests exercise the bias path."""
    rng = np.random.default_rng(seed)
    Ws, bs = [], []
    for n_in, n_out in zip(dims[:-1], dims[1:]):
        Ws.append(truncated_normal((n_in, n_out), np.sqrt(2.0 / n_in), rng))
        bs.append((bias_scale * rng.standard_normal(n_out)).astype(np.float32))
    return Ws, bs


def ml_codebook_model(L: int = 7):
    """'Perfect weights' self-test from SURVEY.md Appendix D: a 1-layer net whose logits are
    y @ C^T with C the +-1 code book of the (133,171) code == brute-force ML decoding."""
    from .encoders import generate_class_bits, encode_133171
    C = 2.0 * encode_133171(generate_class_bits(L)).astype(np.float32) - 1.0  # [2^L, 2L]
    return [C.T.copy()], [np.zeros(C.shape[0], np.float32)], [ACT_IDENTITY]
