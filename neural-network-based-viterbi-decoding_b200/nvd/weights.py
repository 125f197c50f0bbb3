"""Decoder parameters: containers, `.npz` interchange keyed by the reference's TF variable names
(weights_h1 ... bias_out, /root/reference/src/fnn_viterbi.py:436-465) and
the reference's initialisers (truncated normal, stddev = sqrt(2 / n_in), zero biases,
/root/reference/src/fnn_viterbi.py:430-466)."""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np

from . import _lib, tf_bundle

_HERE = os.path.dirname(os.path.abspath(__file__))
WEIGHTS_DIR = os.path.join(os.path.dirname(_HERE), "weights")
DEFAULT_FNN = os.path.join(WEIGHTS_DIR, "fnn_133171_L7.npz")

_ACT = {"identity": _lib.ACT_IDENTITY, "tanh": _lib.ACT_TANH, "relu": _lib.ACT_RELU}


@dataclass
class MlpModel:
    """Ws[l] is [in, out] (x @ W + b); acts[l] in {0 identity, 1 tanh, 2 relu}."""
    Ws: list
    bs: list
    acts: list
    meta: dict = field(default_factory=dict)

    @property
    def dims(self):
        return [self.Ws[0].shape[0]] + [W.shape[1] for W in self.Ws]


@dataclass
class LstmModel:
    """TF-1.x LSTMCell kernels [1+H, 4H] (gate order i, j, f, o), biases [4H], linear [H] -> 1."""
    enc_W: np.ndarray
    enc_b: np.ndarray
    dec_W: np.ndarray
    dec_b: np.ndarray
    lin_w: np.ndarray
    lin_b: np.ndarray

    @property
    def H(self):
        return self.enc_b.shape[0] // 4

    def as_dict(self):
        return dict(enc_W=self.enc_W, enc_b=self.enc_b, dec_W=self.dec_W, dec_b=self.dec_b,
                    lin_w=self.lin_w, lin_b=self.lin_b)


def truncated_normal(shape, stddev, rng: np.random.Generator) -> np.ndarray:
    """tf.truncated_normal: values beyond two standard deviations are redrawn."""
    out = rng.standard_normal(shape)
    bad = np.abs(out) > 2.0
    while bad.any():
        out[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(out) > 2.0
    return (out * stddev).astype(np.float32)


_LAYER_KEYS = ["h1", "h2", "h3", "h4", "h5", "h6", "h7"]
_BIAS_KEYS = ["b1", "b2", "b3", "b4", "b5", "b6", "b7"]


def init_mlp(dims, acts, seed=None) -> MlpModel:
    rng = np.random.default_rng(seed)
    Ws = [truncated_normal((i, o), np.sqrt(2.0 / i), rng) for i, o in zip(dims[:-1], dims[1:])]
    bs = [np.zeros(o, np.float32) for o in dims[1:]]
    return MlpModel(Ws, bs, [_ACT.get(a, a) for a in acts])


def to_dicts(model: MlpModel):
    """-> (weights{'h1',...,'out'}, biases{'b1',...,'out'}) as the reference's initialize_weights returns."""
    n = len(model.Ws)
    w = {_LAYER_KEYS[i]: model.Ws[i] for i in range(n - 1)}
    b = {_BIAS_KEYS[i]: model.bs[i] for i in range(n - 1)}
    w["out"], b["out"] = model.Ws[-1], model.bs[-1]
    return w, b


def from_dicts(weights: dict, biases: dict, acts) -> MlpModel:
    keys = [k for k in _LAYER_KEYS if k in weights]
    Ws = [np.ascontiguousarray(weights[k], dtype=np.float32) for k in keys] + [np.ascontiguousarray(weights["out"], np.float32)]
    bs = [np.ascontiguousarray(biases[_BIAS_KEYS[_LAYER_KEYS.index(k)]], dtype=np.float32) for k in keys]
    bs.append(np.ascontiguousarray(biases["out"], np.float32))
    acts = [_ACT.get(a, a) for a in acts]
    if len(acts) != len(Ws):
        raise ValueError("acts length must match the number of layers")
    return MlpModel(Ws, bs, acts)


def save_mlp(path: str, model: MlpModel) -> None:
    """`.npz` keyed by the TF variable names the reference's Saver would write."""
    w, b = to_dicts(model)
    arrays = {f"weights_{k}": v for k, v in w.items()}
    arrays.update({("bias_" + (k if k == "out" else "h" + k[1:])): v for k, v in b.items()})
    arrays["acts"] = np.asarray(model.acts, dtype=np.int32)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    np.savez(path, **arrays)


def load_mlp(path: str) -> MlpModel:
    """`.npz` written by save_mlp, or — when there is no such file but `<path minus .npz>.index` exists — a TensorFlow
    checkpoint as the reference's tf.train.Saver writes it (nvd/tf_bundle.py; variables weights_h1.., bias_h1..,
    weights_out, bias_out as named in src/fnn_viterbi_mixed_snr.py:533-567; optimizer slots are ignored)."""
    prefix = path[:-4] if path.endswith(".npz") else path
    if not os.path.exists(path) and tf_bundle.is_bundle(prefix):
        z = tf_bundle.read_bundle(prefix)
        files = set(z)
        meta = {"path": prefix, "format": "tf-bundle"}
    else:
        z = np.load(path)
        files = set(z.files)
        meta = {"path": path}
    keys = [k for k in _LAYER_KEYS if f"weights_{k}" in files]
    Ws = [np.asarray(z[f"weights_{k}"], dtype=np.float32) for k in keys] + [np.asarray(z["weights_out"], dtype=np.float32)]
    bs = [np.asarray(z[f"bias_{k}"], dtype=np.float32) for k in keys] + [np.asarray(z["bias_out"], dtype=np.float32)]
    acts = np.asarray(z["acts"]).tolist() if "acts" in files else [_lib.ACT_TANH] * len(Ws)
    return MlpModel(Ws, bs, acts, meta=meta)


def save_tf_checkpoint(prefix: str, model: MlpModel) -> None:
    """The model's variables as a TensorFlow checkpoint (`<prefix>.index` + `.data-00000-of-00001`) under the names the
    reference's graphs give them, so that `saver.restore(sess, prefix)` there finds them."""
    w, b = to_dicts(model)
    tensors = {f"weights_{k}": v for k, v in w.items()}
    tensors.update({("bias_" + (k if k == "out" else "h" + k[1:])): v for k, v in b.items()})
    tf_bundle.write_bundle(prefix, {k: np.asarray(v, dtype=np.float32) for k, v in tensors.items()})


def load_default_fnn() -> MlpModel:
    """Trained 14-128-64-32-128 tanh decoder for the (133,171) L=7 code when the fixture is
    present (tools/train_fnn.py), else seed-fixed reference-initialised weights."""
    if os.path.exists(DEFAULT_FNN):
        return load_mlp(DEFAULT_FNN)
    m = init_mlp([14, 128, 64, 32, 128], ["tanh"] * 4, seed=20170710)
    rng = np.random.default_rng(7)
    m.bs = [(0.05 * rng.standard_normal(b.shape)).astype(np.float32) for b in m.bs]
    return m


FNN_57 = os.path.join(WEIGHTS_DIR, "fnn_57_L7.npz")
BITWISE = os.path.join(WEIGHTS_DIR, "bitwise_133171_L512_bit0.npz")
RNN = os.path.join(WEIGHTS_DIR, "rnn_133171_L100.npz")


def load_fnn_57() -> MlpModel:
    """Trained 14-128-64-32-128 tanh decoder for the reference's (7,5) encoder, quirk included
    (BASELINE.json configs[0]; tools/train_fnn.py --code 57)."""
    return load_mlp(FNN_57)


def load_default_bitwise() -> MlpModel:
    """Per-bit predictor 1036-30-2 (ReLU, ReLU) for bit 0, start state 0 (tools/train_long_frame_decoders.py bitwise),
    else seed-fixed reference-initialised weights."""
    if os.path.exists(BITWISE):
        return load_mlp(BITWISE)
    m = init_mlp([1036, 30, 2], ["relu", "relu"], seed=20170711)
    m.bs = [np.full(b.shape, 0.05, np.float32) for b in m.bs]
    return m


def load_default_lstm() -> LstmModel:
    """LSTM(20) seq2seq for L = 100 (tools/train_long_frame_decoders.py rnn), else seed-fixed Glorot weights."""
    if os.path.exists(RNN):
        z = np.load(RNN)
        return LstmModel(**{k: z[k].astype(np.float32) for k in ("enc_W", "enc_b", "dec_W", "dec_b", "lin_w", "lin_b")})
    return init_lstm(20, seed=20170712)


def init_lstm(H=20, seed=None) -> LstmModel:
    rng = np.random.default_rng(seed)

    def glorot(shape):
        lim = np.sqrt(6.0 / (shape[0] + shape[1]))
        return rng.uniform(-lim, lim, size=shape).astype(np.float32)

    return LstmModel(glorot((1 + H, 4 * H)), np.zeros(4 * H, np.float32), glorot((1 + H, 4 * H)),
                     np.zeros(4 * H, np.float32), glorot((H, 1)).reshape(H), np.zeros(1, np.float32))
