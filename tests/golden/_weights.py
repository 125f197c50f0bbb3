"""Seeded parameter arrays keyed by the TensorFlow variable NAME the reference gives them
(weights_h1 ... bias_out, rnn/lstm_cell/kernel ...).  tools/tf_numpy_shim/tensorflow.py
materialises the reference's tf.Variable calls through these functions when the golden
outputs are generated, and the tests regenerate the identical arrays from the same
(name, shape, seed) — so tests/golden/reference_forward.npz holds inputs and outputs only."""
import zlib

import numpy as np


def _rng(name: str, seed: int) -> np.random.Generator:
    return np.random.default_rng([int(seed), zlib.crc32(name.encode())])


def truncated_normal(name, shape, stddev, seed):
    """tf.truncated_normal: N(0, stddev) with anything beyond 2 stddev redrawn."""
    rng = _rng(name, seed)
    out = rng.standard_normal(shape)
    bad = np.abs(out) > 2.0
    while bad.any():
        out[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(out) > 2.0
    return (out * stddev).astype(np.float32)


def bias(name, shape, scale, seed):
    """Zero (the reference's init) for scale = 0, else scale * N(0,1) so the bias path is exercised."""
    if scale == 0.0:
        return np.zeros(shape, np.float32)
    return (scale * _rng(name, seed).standard_normal(shape)).astype(np.float32)


def uniform(name, shape, limit, seed):
    return _rng(name, seed).uniform(-limit, limit, shape).astype(np.float32)
