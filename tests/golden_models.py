"""Rebuild the parameters behind tests/golden/reference_forward.npz: the arrays that
tools/tf_numpy_shim/tensorflow.py handed to the reference's tf.Variable calls, by variable NAME
(tests/golden/_weights.py), with the shapes and stddevs the reference's own init code asks for
(src/fnn_viterbi.py:430-466, src/fnn_viterbi_bitwise.py:345-384, src/fnn_viterbi_seq_predict.py:365-404)."""
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import _weights  # noqa: E402

ACT_TANH, ACT_RELU = 1, 2
BIAS_SCALE = 0.05


def _mlp(layers, seed):
    """layers: [(weight name, bias name, n_in, n_out, stddev)]"""
    Ws = [_weights.truncated_normal(wn, (a, b), sd, seed) for wn, _, a, b, sd in layers]
    bs = [_weights.bias(bn, (b,), BIAS_SCALE, seed) for _, bn, a, b, sd in layers]
    return Ws, bs


def fnn(seed):
    d = [14, 128, 64, 32, 128]
    names = ["h1", "h2", "h3", "out"]
    Ws, bs = _mlp([(f"weights_{n}", f"bias_{n}", d[i], d[i + 1], math.sqrt(2.0 / d[i])) for i, n in enumerate(names)], seed)
    return Ws, bs, [ACT_TANH] * 4


def bitwise(seed):
    # the graph that is actually wired: 1036 -> 30 -> 2, ReLU on both; Wout is [n_hidden_1, 2] with
    # stddev sqrt(2 / n_hidden_4) (src/fnn_viterbi_bitwise.py:371-377), h2..h4 are created but unused
    Ws, bs = _mlp([("weights_h1", "bias_h1", 1036, 30, math.sqrt(2.0 / 1036)),
                   ("weights_out", "bias_out", 30, 2, math.sqrt(2.0 / 100))], seed)
    return Ws, bs, [ACT_RELU] * 2


def seq_predict(seed):
    d = [200, 500, 500, 500, 500, 2]
    names = ["h1", "h2", "h3", "h4", "out"]
    Ws, bs = _mlp([(f"weights_{n}", f"bias_{n}", d[i], d[i + 1], math.sqrt(2.0 / d[i])) for i, n in enumerate(names)], seed)
    return Ws, bs, [ACT_TANH] * 5


def rnn(seed, H=20):
    lim = math.sqrt(6.0 / (1 + H + 4 * H))
    p = {
        "enc_W": _weights.uniform("rnn/lstm_cell/kernel", (1 + H, 4 * H), lim, seed),
        "enc_b": _weights.bias("rnn/lstm_cell/bias", (4 * H,), BIAS_SCALE, seed),
        "dec_W": _weights.uniform("plain_decoder/lstm_cell/kernel", (1 + H, 4 * H), lim, seed),
        "dec_b": _weights.bias("plain_decoder/lstm_cell/bias", (4 * H,), BIAS_SCALE, seed),
        "lin_w": _weights.uniform("fully_connected/weights", (H, 1), math.sqrt(6.0 / (H + 1)), seed)[:, 0],
        "lin_b": _weights.bias("fully_connected/biases", (1,), BIAS_SCALE, seed),
    }
    return p
