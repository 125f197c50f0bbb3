"""Drop-in for /root/reference/src/fnn_viterbi.py — same names, signatures, array conventions and
printed lines; the work is done by libnvd's CUDA kernels through ctypes (nvd.*).

The reference builds TensorFlow-1.x graph nodes; without TF the graph functions are EAGER:
`initialize_weights` returns dicts of numpy fp32 arrays with the reference's keys / shapes / init
distribution, `multilayer_perceptron(x, weights, biases)` takes a numpy [B, n_input] array and
returns the [B, n_classes] soft outputs computed on the GPU, `run_network` trains (torch autograd,
off the hot path) and decodes the test set with the CUDA decoder.  There is no CPU fallback.
"""
from __future__ import annotations

from random import randint
from typing import Dict, Optional, Tuple

import numpy as np

from nvd import runtime, training, weights as _weights
from nvd.montecarlo import class_to_packed
from nvd import _lib

ENCODER = "133171"   # the reference's dataset builder hard-wires encode_133171 (fnn_viterbi.py:352)


class DecoderConfig:
    """Attribute bag identical to /root/reference/src/fnn_viterbi.py:27-92 (+ `encoder`, `precision`)."""

    def __init__(self, block_length: int = 7, rate: int = 2, snr_db: float = 2.0, n_hidden_1: int = 128,
                 n_hidden_2: int = 64, n_hidden_3: int = 32, learning_rate: float = 0.0001, training_epochs: int = 200,
                 batch_size: int = 256, beta_regularization: float = 0.001, samples_per_class_train: int = 500,
                 samples_per_class_test: int = 10000, encoder: str = ENCODER, precision: Optional[str] = None):
        self.block_length = block_length
        self.rate = rate
        self.all_classes = 2 ** block_length
        self.num_classes = self.all_classes
        self.n_input = rate * block_length
        self.n_hidden_1, self.n_hidden_2, self.n_hidden_3 = n_hidden_1, n_hidden_2, n_hidden_3
        self.n_classes = self.all_classes
        self.learning_rate = learning_rate
        self.training_epochs = training_epochs
        self.batch_size = batch_size
        self.beta_regularization = beta_regularization
        self.display_step = 1
        self.num_training_samples = self.num_classes * samples_per_class_train
        self.samples_per_class_train = samples_per_class_train
        self.num_test_samples = self.num_classes * samples_per_class_test
        self.samples_per_class_test = samples_per_class_test
        self.snr_db = snr_db
        self.snr = 10 ** (snr_db / 10.0)
        self.n0 = 1.0 / self.snr
        self.sigma = np.sqrt(self.n0 * 0.5)
        self.encoder = encoder
        self.precision = precision


# ---- source / encoders (K1) ---------------------------------------------------------------------
def generate_info_sequence(block_length: int) -> np.ndarray:
    info_seq = np.zeros(block_length, dtype=np.int32)
    for i in range(block_length):
        info_seq[i] = randint(0, 1)
    return info_seq


def encode_57(info_seq: np.ndarray) -> np.ndarray:
    """(7,5) encoder exactly as the reference computes it (g=5 emitted first, c[1] = b0^b1)."""
    return runtime.get_sim().encode_bits(np.asarray(info_seq), "57")


def encode_133171(bits: np.ndarray) -> np.ndarray:
    """(133,171) K=7 encoder, zero state, unterminated, interleaved [g133(i), g171(i)]."""
    return runtime.get_sim().encode_bits(np.asarray(bits), "133171")


encode133171 = encode_133171   # alias imported by production_wrapper.py:32-38


# ---- channel (K2) ---------------------------------------------------------------------------------
def modulate_awgn(encoded_seq: np.ndarray, sigma: float) -> np.ndarray:
    """BPSK (0 -> -1, 1 -> +1) + N(0, sigma^2).  Noise is Philox-keyed by a process-wide frame
    counter (nvd.runtime.seed(s) makes a run reproducible); returned as float64 like the reference."""
    enc = np.asarray(encoded_seq)
    rows = enc.reshape(-1, enc.shape[-1])
    f0 = runtime.take_frames(rows.shape[0])
    y = runtime.get_sim().channel_bits(rows != 0, float(sigma), 0, f0, "f32").astype(np.float64)
    return y.reshape(enc.shape)


modulateAwgn = modulate_awgn   # alias imported by production_wrapper.py


# ---- labels --------------------------------------------------------------------------------------------
def generate_class_bits(block_length: int) -> np.ndarray:
    idx = np.arange(2 ** block_length, dtype=np.int64)[:, None]
    sh = np.arange(block_length - 1, -1, -1, dtype=np.int64)[None, :]
    return ((idx >> sh) & 1).astype(np.uint8)


def create_one_hot_labels(labels: np.ndarray, num_classes: int) -> np.ndarray:
    labels = np.asarray(labels).astype(np.int64)
    one_hot = np.zeros((len(labels), num_classes), dtype=np.float32)
    one_hot[np.arange(len(labels)), labels] = 1.0
    return one_hot


def _noisy_class_major(config, samples_per_class: int, sigma: float) -> Tuple[np.ndarray, np.ndarray]:
    """`samples_per_class` noisy copies of every code word, class-major (fnn_viterbi.py:369-388),
    built by K1 + K2 on the GPU."""
    sim = runtime.get_sim()
    L = config.block_length
    labels = np.repeat(np.arange(config.num_classes, dtype=np.int64), samples_per_class)
    B = len(labels)
    nbytes = (L + 7) // 8
    words = class_to_packed(labels, L)
    info = np.stack([(words >> (8 * b)) & 0xFF for b in range(nbytes)], axis=1).astype(np.uint8)
    d_info = sim.to_device(info)
    d_coded = sim.encode(getattr(config, "encoder", ENCODER), d_info, L, B)
    f0 = runtime.take_frames(B)
    d_y = sim.channel(d_coded, 2 * L, B, float(sigma), 0, f0, "f32")
    y = d_y.download(np.float32, (B, 2 * L))
    for b in (d_info, d_coded, d_y):
        sim.free(b)
    return y, labels.astype(np.int32)


def train_test_dataset(config: Optional[DecoderConfig] = None):
    """-> (training_set, one_hot_training_labels, test_set, one_hot_test_labels, class_bits).
    Zero-argument form (production_wrapper.py:98) uses the default configuration."""
    config = config or DecoderConfig()
    class_bits = generate_class_bits(config.block_length)
    training_set, training_labels = _noisy_class_major(config, config.samples_per_class_train, config.sigma)
    test_set, test_labels = _noisy_class_major(config, config.samples_per_class_test, config.sigma)
    return (training_set, create_one_hot_labels(training_labels, config.n_classes), test_set,
            create_one_hot_labels(test_labels, config.n_classes), class_bits)


# ---- decoder (K3) ----------------------------------------------------------------------------------------
def initialize_weights(config: Optional[DecoderConfig] = None) -> Tuple[Dict, Dict]:
    """Truncated-normal (stddev = sqrt(2/n_in)) weights [in, out] and zero biases, keys
    {'h1','h2','h3','out'} / {'b1','b2','b3','out'} (fnn_viterbi.py:430-466)."""
    config = config or DecoderConfig()
    dims = [config.n_input, config.n_hidden_1, config.n_hidden_2, config.n_hidden_3, config.n_classes]
    return _weights.to_dicts(_weights.init_mlp(dims, ["tanh"] * 4))


get_weights = initialize_weights   # alias imported by production_wrapper.py


def _model_from(weights: Dict, biases: Dict) -> _weights.MlpModel:
    return _weights.from_dicts(weights, biases, ["tanh"] * (len(weights)))


def multilayer_perceptron(x, weights: Dict, biases: Dict, precision: Optional[str] = None) -> np.ndarray:
    """out = tanh(tanh(tanh(tanh(x W1 + b1) W2 + b2) W3 + b3) Wout + bout) — tanh on the output too
    (fnn_viterbi.py:495-508) — evaluated on the GPU; returns numpy [B, n_classes]."""
    sim = runtime.get_sim()
    h = runtime.cached_mlp(_model_from(weights, biases))      # loaded once per distinct parameter set
    _, soft = sim.decode(h, np.asarray(x, dtype=np.float32), _pick_precision(sim, h, precision))
    return soft


def _pick_precision(sim, h, precision):
    p = precision or runtime.precision()
    if p == "bf16" and not sim.mlp_tc_supported(h):
        raise _lib.NvdError(-4, "bf16 tensor-core path does not cover this layer shape; use precision='fp32'")
    return p


def predict(x, weights: Dict, biases: Dict, precision: Optional[str] = None) -> np.ndarray:
    """argmax(multilayer_perceptron(x), 1) — the decode of fnn_viterbi.py:606-610 — as int64 [B]."""
    sim = runtime.get_sim()
    h = runtime.cached_mlp(_model_from(weights, biases))
    pred, _ = sim.decode(h, np.asarray(x, dtype=np.float32), _pick_precision(sim, h, precision), want_soft=False)
    return pred.astype(np.int64)


def run_network(config: DecoderConfig, training_set, training_labels, test_set, test_labels,
                weights: Optional[Dict] = None, biases: Optional[Dict] = None) -> np.ndarray:
    """Train (Adam, softmax-CE on the tanh outputs + beta * L2(h1,h2,h3), fnn_viterbi.py:549-603), then
    decode the whole test set on the GPU; prints the reference's lines and returns the predictions."""
    if weights is None or biases is None:
        weights, biases = initialize_weights(config)
    model = _model_from(weights, biases)
    training.train_mlp(model, training_set, training_labels, acts=model.acts, loss="softmax_ce",
                       beta=config.beta_regularization, l2_layers=(0, 1, 2), lr=config.learning_rate,
                       epochs=config.training_epochs, batch_size=config.batch_size, display_step=config.display_step)
    print("Optimization Finished!")
    w, b = _weights.to_dicts(model)
    weights.update(w)
    biases.update(b)
    predictions = predict(test_set, weights, biases, getattr(config, "precision", None))
    accuracy = float(np.mean(predictions == np.argmax(test_labels, axis=1)))
    print(f"Accuracy: {accuracy:.4f}")
    return predictions


# ---- metrics (K5) ------------------------------------------------------------------------------------------
def error_counts(labels_idx: np.ndarray, predictions: np.ndarray, block_length: int) -> Tuple[int, int]:
    """(bit_errors, frame_errors) by the fused popcount reduction on the GPU."""
    sim = runtime.get_sim()
    B = len(predictions)
    stride = 1 if block_length <= 8 else 2 if block_length <= 16 else 4
    dt = {1: np.uint8, 2: np.uint16, 4: np.uint32}[stride]
    d_pred = sim.to_device(np.asarray(predictions).astype(dt))
    d_lab = sim.to_device(class_to_packed(labels_idx, block_length).astype(dt))
    counters = sim.zeros(32)
    sim.count_errors(d_pred, d_lab, stride, block_length, B, _lib.COUNT_PRED_IS_MSB_CLASS, counters)
    c = counters.download(np.uint64, (4,))
    for buf in (d_pred, d_lab, counters):
        sim.free(buf)
    return int(c[0]), int(c[1])


def calculate_error(test_labels: np.ndarray, predictions: np.ndarray, class_bits: np.ndarray, config: DecoderConfig) -> None:
    """Prints Accuracy / Frame Error Rate / Bit Error Rate / Total Bit Errors (fnn_viterbi.py:643-674).
    The reference's BER loop assumes class-major order (:655-667); labels here come from the one-hot
    rows, which is identical for class-major sets and also correct for shuffled ones."""
    test_labels_idx = np.argmax(test_labels, axis=1).astype(int)
    bit_errors, frame_errors = error_counts(test_labels_idx, predictions, config.block_length)
    fer = frame_errors / len(predictions)
    accuracy = 1.0 - fer
    print(f"Accuracy: {accuracy:.4f}")
    print(f"Frame Error Rate: {fer:.6f}")
    total_bits = len(predictions) * config.block_length
    ber = bit_errors / total_bits
    print(f"Bit Error Rate: {ber:.6f}")
    print(f"Total Bit Errors: {bit_errors}/{total_bits}")


def main() -> None:
    print("=" * 70)
    print("Neural Network-Based Viterbi Decoder")
    print("=" * 70)
    config = DecoderConfig(block_length=7, rate=2, snr_db=2.0, n_hidden_1=128, n_hidden_2=64, n_hidden_3=32,
                           learning_rate=0.0001, training_epochs=200, batch_size=256, beta_regularization=0.001,
                           samples_per_class_train=500, samples_per_class_test=10000)
    print("\nConfiguration:")
    print(f"  Block Length: {config.block_length}")
    print(f"  SNR: {config.snr_db} dB")
    print(f"  Number of Classes: {config.num_classes}")
    print(f"  Training Samples: {config.num_training_samples}")
    print(f"  Test Samples: {config.num_test_samples}")
    print(f"  Network: {config.n_input} -> {config.n_hidden_1} -> {config.n_hidden_2} -> {config.n_hidden_3} -> {config.n_classes}")
    print("\nGenerating datasets...")
    training_set, training_labels, test_set, test_labels, class_bits = train_test_dataset(config)
    print(f"Training set shape: {training_set.shape}")
    print(f"Test set shape: {test_set.shape}")
    print("\nTraining neural network...")
    predictions = run_network(config, training_set, training_labels, test_set, test_labels)
    print("\nEvaluating performance...")
    calculate_error(test_labels, predictions, class_bits, config)
    print("\n" + "=" * 70)
    print("Decoding complete!")
    print("=" * 70)


if __name__ == "__main__":
    main()
