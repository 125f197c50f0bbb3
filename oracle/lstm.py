"""Oracle: LSTM seq2seq decoder (TEST INFRASTRUCTURE, parity UNPINNED by the reference).

Restates build_seq2seq_model (/root/reference/src/rnn_viterbi.py:235-325):
  encoder  tf.contrib.rnn.LSTMCell(H) over encoder_inputs [T, B, 1], zero state   :278-286
  decoder  LSTMCell(H) over decoder_inputs [T, B, 1], initial state = encoder's    :289-299
  logits   tf.contrib.layers.linear(decoder_outputs, 1)                            :302-305
TF 1.x LSTMCell (third-party, absent): kernel [1+H, 4H] applied to concat([x, h]);
gate order i, j, f, o; forget_bias = 1.0; no peepholes / projection / clipping:
  c' = sigmoid(f + 1) * c + sigmoid(i) * tanh(j);   h' = sigmoid(o) * tanh(c')
prepare_batch (:328-377): decoder_inputs = [EOS, info(+-1)..., 0 x (L-1)],
decoder_targets = [info(+-1)..., EOS, 0 x (L-1)], time-major.
"""
from __future__ import annotations

import numpy as np


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def lstm_cell_step(x, c, h, W, b, forget_bias=1.0):
    """x [B, I]; c, h [B, H]; W [I+H, 4H]; b [4H]."""
    z = np.concatenate([x, h], axis=1) @ W + b
    i, j, f, o = np.split(z, 4, axis=1)
    c_new = _sigmoid(f + forget_bias) * c + _sigmoid(i) * np.tanh(j)
    h_new = _sigmoid(o) * np.tanh(c_new)
    return c_new, h_new


def seq2seq_forward(enc_in, dec_in, params, dtype=np.float64, autoregressive=False,
                    eos=1.0):
    """enc_in [B, T_enc], dec_in [B, T_dec] (ignored if autoregressive) -> logits [B, T_dec].

    params: dict enc_W [1+H,4H], enc_b [4H], dec_W, dec_b, lin_w [H], lin_b scalar.
    Autoregressive mode (not in the reference, which teacher-forces even at test time,
    rnn_viterbi.py:362-371,464-469) feeds sign(previous logit) back, starting from EOS."""
    enc_in = np.asarray(enc_in, dtype=dtype)
    B, T_enc = enc_in.shape
    p = {k: np.asarray(v, dtype=dtype) for k, v in params.items()}
    H = p["enc_b"].shape[0] // 4
    c = np.zeros((B, H), dtype=dtype)
    h = np.zeros((B, H), dtype=dtype)
    for t in range(T_enc):
        c, h = lstm_cell_step(enc_in[:, t:t + 1], c, h, p["enc_W"], p["enc_b"])
    T_dec = np.asarray(dec_in).shape[1] if dec_in is not None else T_enc
    logits = np.zeros((B, T_dec), dtype=dtype)
    prev = np.full((B, 1), eos, dtype=dtype)
    if not autoregressive:
        dec_in = np.asarray(dec_in, dtype=dtype)
    for t in range(T_dec):
        x = prev if autoregressive else dec_in[:, t:t + 1]
        c, h = lstm_cell_step(x, c, h, p["dec_W"], p["dec_b"])
        lg = h @ p["lin_w"].reshape(H, 1) + p["lin_b"]
        logits[:, t] = lg[:, 0]
        if autoregressive:
            prev = np.where(lg > 0, 1.0, -1.0).astype(dtype)
    return logits


def prepare_batch(info_pm1, encoded, block_length, eos=1.0):
    """Batch-major statement of rnn_viterbi.py:328-377 -> (enc_in [B,2L], dec_in [B,2L], dec_tgt [B,2L])."""
    info_pm1 = np.asarray(info_pm1, dtype=np.float64)
    B = info_pm1.shape[0]
    L = block_length
    dec_in = np.concatenate([np.full((B, 1), eos), info_pm1, np.zeros((B, L - 1))], axis=1)
    dec_tgt = np.concatenate([info_pm1, np.full((B, 1), eos), np.zeros((B, L - 1))], axis=1)
    return np.asarray(encoded, dtype=np.float64), dec_in, dec_tgt


def init_params(H=20, seed=0, scale=None):
    """Seed-fixed LSTM parameters (TF default glorot-uniform kernels, zero biases)."""
    rng = np.random.default_rng(seed)

    def glorot(shape):
        lim = np.sqrt(6.0 / (shape[0] + shape[1])) if scale is None else scale
        return rng.uniform(-lim, lim, size=shape).astype(np.float32)

    return {
        "enc_W": glorot((1 + H, 4 * H)), "enc_b": np.zeros(4 * H, np.float32),
        "dec_W": glorot((1 + H, 4 * H)), "dec_b": np.zeros(4 * H, np.float32),
        "lin_w": glorot((H, 1)).reshape(H), "lin_b": np.zeros(1, np.float32),
    }
