"""A numpy-backed stand-in for the handful of TensorFlow-1.x symbols the reference's graph
functions touch, so that the REFERENCE'S OWN wiring (multilayer_perceptron, build_network,
build_seq2seq_model under /root/reference/src) runs unchanged and eagerly in this container,
where TensorFlow is not installable.  Used only by tools/gen_golden_forward.py to produce
tests/golden/reference_forward.npz; nothing in the product or the tests imports it.

What it is:  tf.matmul / tf.add / tf.nn.tanh / tf.nn.relu / tf.argmax ... are the numpy
functions of the same meaning, in float32 like TF's CPU kernels.  tf.Variable(init, name=...)
materialises a seeded array chosen by the variable's NAME (tests regenerate the same arrays
from tests/golden/_weights.py), so no weight file has to be stored.
What it is not:  TensorFlow.  For the MLPs every op is elementary and the shim adds nothing
of its own.  For the LSTM the cell arithmetic below is a restatement of TF-1.x LSTMCell
(gate order i, j, f, o; forget_bias 1.0), so rnn_viterbi's golden pins the reference's
WIRING (time-major feeds, encoder state -> decoder, linear head), not TensorFlow's cell.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "tests", "golden"))
import _weights  # noqa: E402  (tests/golden/_weights.py: seeded arrays by variable name)

float32 = np.float32
int64 = np.int64

# -- state the generating script sets / reads ---------------------------------------------------
SEED = 0                 # seed of every materialised variable
BIAS_SCALE = 0.0         # tf.zeros biases become BIAS_SCALE * N(0,1) (0 = the reference's zero init)
FEEDS = {}               # placeholder name -> array
VARIABLES = {}           # name -> materialised array (in creation order)
TAPS = {}                # named intermediate results (e.g. the linear head's logits)


def reset(seed=0, bias_scale=0.0, feeds=None):
    global SEED, BIAS_SCALE
    SEED, BIAS_SCALE = seed, bias_scale
    FEEDS.clear()
    FEEDS.update(feeds or {})
    VARIABLES.clear()
    TAPS.clear()


class _Init:
    def __init__(self, kind, shape, stddev=0.0):
        self.kind, self.shape, self.stddev = kind, tuple(int(s) for s in shape), float(stddev)


def truncated_normal(shape, stddev=1.0, **_):
    return _Init("truncated_normal", shape, stddev)


def zeros(shape, **_):
    return _Init("zeros", shape)


class Tensor(np.ndarray):
    pass


class Variable(np.ndarray):
    def __new__(cls, initial_value, name=None, **_):
        if isinstance(initial_value, _Init):
            if initial_value.kind == "truncated_normal":
                v = _weights.truncated_normal(name, initial_value.shape, initial_value.stddev, SEED)
            else:
                v = _weights.bias(name, initial_value.shape, BIAS_SCALE, SEED)
        else:
            v = np.asarray(initial_value, dtype=np.float32)
        VARIABLES[name] = v
        return v


def placeholder(dtype=None, shape=None, name=None):
    return np.asarray(FEEDS[name], dtype=np.float32)


def matmul(a, b):
    return np.matmul(np.asarray(a, np.float32), np.asarray(b, np.float32))


def add(a, b):
    return np.add(a, b, dtype=np.float32)


def expand_dims(x, axis):
    return np.expand_dims(x, axis)


def argmax(x, axis=None):
    return np.argmax(x, axis=axis)


def reduce_mean(x):
    return np.mean(x)


def squared_difference(a, b):
    return np.square(np.asarray(a, np.float32) - np.asarray(b, np.float32))


class nn:
    @staticmethod
    def tanh(x):
        return np.tanh(np.asarray(x, np.float32))

    @staticmethod
    def relu(x):
        return np.maximum(np.asarray(x, np.float32), np.float32(0))

    @staticmethod
    def dropout(x, keep_prob):
        assert float(keep_prob) == 1.0, "golden generation runs the inference graph (keep_prob = 1)"
        return x

    @staticmethod
    def dynamic_rnn(cell, inputs, dtype=None, time_major=False, initial_state=None, scope=None):
        assert time_major
        scope = scope or "rnn"
        T, B, _ = inputs.shape
        c, h = initial_state if initial_state is not None else (np.zeros((B, cell.H), np.float32),) * 2
        outs = []
        for t in range(T):
            c, h = cell.step(inputs[t], c, h, scope)
            outs.append(h)
        return np.stack(outs), (c, h)


def _sigmoid(x):
    return (1.0 / (1.0 + np.exp(-x.astype(np.float64)))).astype(np.float32)


class _LSTMCell:
    """TF-1.x tf.contrib.rnn.LSTMCell(num_units) defaults: kernel [input+H, 4H] on concat([x, h]),
    gates i, j, f, o, forget_bias 1.0, no peepholes / projection / clipping."""

    def __init__(self, num_units):
        self.H = int(num_units)

    def step(self, x, c, h, scope):
        kname, bname = f"{scope}/lstm_cell/kernel", f"{scope}/lstm_cell/bias"
        if kname not in VARIABLES:
            n_in = x.shape[1] + self.H
            VARIABLES[kname] = _weights.uniform(kname, (n_in, 4 * self.H), float(np.sqrt(6.0 / (n_in + 4 * self.H))), SEED)
            VARIABLES[bname] = _weights.bias(bname, (4 * self.H,), BIAS_SCALE, SEED)
        z = np.concatenate([x, h], axis=1).astype(np.float32) @ VARIABLES[kname] + VARIABLES[bname]
        i, j, f, o = np.split(z, 4, axis=1)
        c_new = _sigmoid(f + np.float32(1.0)) * c + _sigmoid(i) * np.tanh(j)
        h_new = _sigmoid(o) * np.tanh(c_new)
        return c_new.astype(np.float32), h_new.astype(np.float32)


class contrib:
    class rnn:
        LSTMCell = _LSTMCell

    class layers:
        @staticmethod
        def linear(inputs, num_outputs):
            H = inputs.shape[-1]
            if "fully_connected/weights" not in VARIABLES:
                VARIABLES["fully_connected/weights"] = _weights.uniform("fully_connected/weights", (H, num_outputs),
                                                                        float(np.sqrt(6.0 / (H + num_outputs))), SEED)
                VARIABLES["fully_connected/biases"] = _weights.bias("fully_connected/biases", (num_outputs,), BIAS_SCALE, SEED)
            out = np.asarray(inputs, np.float32) @ VARIABLES["fully_connected/weights"] + VARIABLES["fully_connected/biases"]
            TAPS["decoder_logits"] = out
            return out


class train:
    class AdamOptimizer:
        def __init__(self, *a, **k):
            pass

        def minimize(self, loss):
            return None
