"""Drop-in for /root/reference/src/fnn_viterbi_mixed_snr.py: the same 14-128-64-32-128 tanh decoder
trained on a range of SNRs and tested at one of them, with the "decode noisy blocks with an old
model" surface (create_model_graph(train_mode=False, model_load_path=...), README.md:33-40).
Names, signatures, return values and printed lines follow the reference; encoder, channel,
decoder forward, argmax and BER counting run in libnvd's CUDA kernels (no CPU fallback).
Models are saved / restored as `.npz` keyed by the reference's TF variable names."""
from __future__ import annotations

import os
from typing import Dict, Optional, Tuple

import numpy as np

import fnn_viterbi as _fv
from fnn_viterbi import (encode_57, encode_133171, generate_class_bits, modulate_awgn,  # noqa: F401
                         multilayer_perceptron, predict, error_counts)
from nvd import runtime, training, weights as _weights

_LAST_MODEL = {"weights": None, "biases": None}


class MixedSNRDecoderConfig:
    """Attribute bag identical to /root/reference/src/fnn_viterbi_mixed_snr.py:33-123."""

    def __init__(self, block_length: int = 7, rate: int = 2, n_hidden_1: int = 128, n_hidden_2: int = 64,
                 n_hidden_3: int = 32, learning_rate: float = 0.0001, training_epochs: int = 300, batch_size: int = 256,
                 beta_regularization: float = 0.001, samples_per_class_train: int = 1000,
                 samples_per_class_test: int = 10000, snr_db_range: Tuple[int, int] = (-2, 10), test_snr_index: int = 0,
                 use_dropout: bool = False, dropout_keep_prob: float = 0.5, use_random_test_set: bool = False,
                 random_test_set_size: int = 5000, encoder: str = "133171", precision: Optional[str] = None):
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
        self.use_dropout = use_dropout
        self.dropout_keep_prob = dropout_keep_prob
        self.samples_per_class_train = samples_per_class_train
        self.samples_per_class_test = samples_per_class_test
        self.num_training_samples = self.num_classes * samples_per_class_train
        self.num_test_samples = self.num_classes * samples_per_class_test
        snr_db_vec = np.arange(snr_db_range[0], snr_db_range[1])
        self.snr_db_vec = snr_db_vec
        self.snr_vec = 10 ** (snr_db_vec / 10.0)
        self.n0_vec = 1.0 / self.snr_vec
        self.sigma_vec = np.sqrt(self.n0_vec * 0.5)
        self.test_snr_index = test_snr_index
        self.samples_per_snr = int(samples_per_class_train / len(snr_db_vec))
        self.use_random_test_set = use_random_test_set
        self.random_test_set_size = random_test_set_size
        self.random_test_length = block_length * random_test_set_size
        self.encoder = encoder
        self.precision = precision


# ---- utilities (host integer helpers; the reference's quirks are kept) ----
def binary_to_decimal(binary_array: np.ndarray) -> int:
    """LSB-first weights 2**arange(n), exactly as the reference code computes it (:277) although its
    docstring says MSB first; `binary_to_decimal_msb` is the documented behaviour."""
    binary_array = np.asarray(binary_array)
    return int(binary_array.dot(2 ** np.arange(binary_array.shape[0])))


def binary_to_decimal_msb(binary_array: np.ndarray) -> int:
    binary_array = np.asarray(binary_array)
    return int(binary_array.dot(2 ** np.arange(binary_array.shape[0])[::-1]))


def extract_bits_single(decimal_value: int, block_length: int) -> np.ndarray:
    bits = np.zeros(block_length, dtype=np.uint8)
    binary_str = format(int(decimal_value), f"0{block_length}b")
    bits[: len(binary_str)] = [int(b) for b in binary_str][:block_length]
    return bits


def calculate_hamming_distance_decimal(val1: int, val2: int, block_length: int) -> int:
    return bin(int(val1) ^ int(val2)).count("1")


def create_one_hot_labels(labels: np.ndarray, num_classes: int) -> np.ndarray:
    labels = np.asarray(labels).astype(np.int64)
    one_hot = np.zeros((len(labels), num_classes), dtype=np.float16)      # float16 here (:371)
    one_hot[np.arange(len(labels)), labels] = 1.0
    return one_hot


# ---- dataset ---------------------------------------------------------------------------------------
def get_dataset(config: MixedSNRDecoderConfig):
    """-> (training_set, test_set, training_labels_onehot, random_test_labels_onehot, random_test_set,
    class_bits), built by the K1/K2 kernels.  Layout follows fnn_viterbi_mixed_snr.py:430-494: per
    class, samples_per_snr rows at each training SNR (the remaining rows of the class stay zero, B-10);
    test rows class-major at sigma_vec[test_snr_index], then shuffled.  The shuffle permutation is
    kept in `config.test_permutation` / `config.test_labels` (the reference drops the test labels and
    rebuilds them unshuffled, B-9)."""
    sim = runtime.get_sim()
    L = config.block_length
    class_bits = generate_class_bits(L)
    book = sim.encode_bits(class_bits, config.encoder)                               # [2^L, 2L]
    n_snr = len(config.snr_db_vec)
    training_set = np.zeros((config.num_training_samples, 2 * L), dtype=np.float32)
    training_labels = np.repeat(np.arange(config.num_classes, dtype=np.int32), config.samples_per_class_train)
    for snr_idx in range(n_snr):
        rows = (np.arange(config.num_classes)[:, None] * config.samples_per_class_train
                + config.samples_per_snr * snr_idx + np.arange(config.samples_per_snr)[None, :]).reshape(-1)
        coded = np.repeat(book, config.samples_per_snr, axis=0)
        f0 = runtime.take_frames(len(rows))
        training_set[rows] = sim.channel_bits(coded, float(config.sigma_vec[snr_idx]), snr_idx, f0, "f32")
    test_sigma = float(config.sigma_vec[config.test_snr_index])
    test_labels = np.repeat(np.arange(config.num_classes, dtype=np.int32), config.samples_per_class_test)
    f0 = runtime.take_frames(len(test_labels))
    test_set = sim.channel_bits(book[test_labels], test_sigma, config.test_snr_index, f0, "f32")
    # random stream test set: one long stream encoded ACROSS block boundaries, labelled LSB-first (B-8)
    random_test_bits = np.random.randint(2, size=config.random_test_length)
    stream = sim.encode_bits(random_test_bits, config.encoder)
    f0 = runtime.take_frames(1)
    random_test_set = sim.channel_bits(stream, test_sigma, config.test_snr_index, f0, "f32").reshape(
        config.random_test_set_size, 2 * L)
    rb = random_test_bits.reshape(config.random_test_set_size, L)
    random_labels = rb.dot(2 ** np.arange(L)).astype(np.uint16)
    perm = np.random.permutation(len(test_labels))
    test_set, test_labels = test_set[perm], test_labels[perm]
    config.test_permutation, config.test_labels = perm, test_labels
    return (training_set, test_set, create_one_hot_labels(training_labels, config.n_classes),
            create_one_hot_labels(random_labels, config.n_classes), random_test_set, class_bits)


# ---- model -----------------------------------------------------------------------------------------
def initialize_weights(config: MixedSNRDecoderConfig) -> Tuple[Dict, Dict]:
    return _fv.initialize_weights(config)


def multilayer_perceptron_with_dropout(x, weights: Dict, biases: Dict, keep_prob: float, rng=None) -> np.ndarray:
    """The dropout twin (:596-627): tanh layers with tf.nn.dropout(keep_prob) after each of the three hidden layers.
    keep_prob = 1 (inference) is multilayer_perceptron itself; for keep_prob < 1 every layer still runs on the GPU
    (one single-layer model each, cached) and the inverted-dropout mask — keep with probability keep_prob, scale
    the kept units by 1 / keep_prob, as tf.nn.dropout does — is drawn on the host from `rng` (numpy Generator)."""
    keep_prob = float(keep_prob)
    if keep_prob >= 1.0:
        return multilayer_perceptron(x, weights, biases)
    if not 0.0 < keep_prob < 1.0:
        raise ValueError("keep_prob must be in (0, 1]")
    rng = rng if rng is not None else np.random.default_rng()
    from nvd import _lib
    sim = runtime.get_sim()
    h = np.asarray(x, dtype=np.float32)
    for wk, bk, last in (("h1", "b1", False), ("h2", "b2", False), ("h3", "b3", False), ("out", "out", True)):
        layer = _weights.MlpModel([np.asarray(weights[wk], np.float32)], [np.asarray(biases[bk], np.float32)], [_lib.ACT_TANH])
        _, h = sim.decode(runtime.cached_mlp(layer), h, "fp32")
        if not last:
            h = h * (rng.random(h.shape) < keep_prob).astype(np.float32) / np.float32(keep_prob)
    return h


def calculate_ber(predictions: np.ndarray, labels: np.ndarray, block_length: int) -> float:
    """sum popcount(pred ^ label) / (N * L) on the GPU reduction kernel (:634-655)."""
    bit_errors, _ = error_counts(np.asarray(labels).astype(np.int64), np.asarray(predictions).astype(np.int64), block_length)
    return bit_errors / (len(predictions) * block_length)


def calculate_accuracy(predictions: np.ndarray, labels: np.ndarray) -> float:
    predictions = np.asarray(predictions)
    return float(np.sum(predictions == np.asarray(labels).astype(int)) / len(predictions))


def save_model(path: str, weights: Dict, biases: Dict) -> None:
    _weights.save_mlp(path if path.endswith(".npz") else path + ".npz", _weights.from_dicts(weights, biases, ["tanh"] * 4))


def load_model(path: str) -> Tuple[Dict, Dict]:
    return _weights.to_dicts(_weights.load_mlp(path if path.endswith(".npz") else path + ".npz"))


def create_model_graph(config: Optional[MixedSNRDecoderConfig] = None, training_set=None, training_labels=None,
                       test_set=None, test_labels=None, random_test_set=None, random_test_labels=None,
                       train_mode: bool = True, model_save_path: Optional[str] = None,
                       model_load_path: Optional[str] = None):
    """Train-or-restore, then decode the test (or random test) set and print accuracy + BER, as
    fnn_viterbi_mixed_snr.py:676-845.  Zero-argument form (README.md:33-40) returns a model handle
    for run_graph()."""
    if config is None:
        config = MixedSNRDecoderConfig()
        weights, biases = initialize_weights(config)
        return {"config": config, "weights": weights, "biases": biases}
    if train_mode:
        weights, biases = initialize_weights(config)
        model = _weights.from_dicts(weights, biases, ["tanh"] * 4)
        print("Starting training...")
        print(f"Training samples: {config.num_training_samples}")
        print(f"SNR values: {config.snr_db_vec} dB")
        print(f"Test SNR: {config.snr_db_vec[config.test_snr_index]} dB")
        training.train_mlp(model, training_set, np.asarray(training_labels, dtype=np.float32), acts=model.acts,
                           loss="softmax_ce", beta=config.beta_regularization, l2_layers=(0, 1, 2), lr=config.learning_rate,
                           epochs=config.training_epochs, batch_size=config.batch_size,
                           dropout_keep=config.dropout_keep_prob if config.use_dropout else None,
                           display_step=config.display_step)
        print("Optimization Finished!")
        weights, biases = _weights.to_dicts(model)
        if model_save_path:
            os.makedirs(os.path.dirname(os.path.abspath(model_save_path)), exist_ok=True)
            save_model(model_save_path, weights, biases)
            print(f"Model saved to {model_save_path}")
        predictions = predict(test_set, weights, biases, config.precision)
        labels = np.argmax(test_labels, axis=1)
        print(f"\nTest Set Accuracy: {calculate_accuracy(predictions, labels):.4f}")
        print(f"Test Set BER: {calculate_ber(predictions, labels, config.block_length):.6f}")
    else:
        if model_load_path:
            weights, biases = load_model(model_load_path)
            print(f"Model loaded from {model_load_path}")
        elif _LAST_MODEL["weights"] is not None:
            weights, biases = _LAST_MODEL["weights"], _LAST_MODEL["biases"]
        else:
            weights, biases = initialize_weights(config)
        if config.use_random_test_set and random_test_set is not None:
            predictions = predict(random_test_set, weights, biases, config.precision)
            labels = np.argmax(random_test_labels, axis=1)
            print(f"Random Test Set Accuracy: {calculate_accuracy(predictions, labels):.4f}")
            print(f"Random Test Set BER: {calculate_ber(predictions, labels, config.block_length):.6f}")
        else:
            predictions = predict(test_set, weights, biases, config.precision)
            labels = np.argmax(test_labels, axis=1)
            print(f"Test Set Accuracy: {calculate_accuracy(predictions, labels):.4f}")
            print(f"Test Set BER: {calculate_ber(predictions, labels, config.block_length):.6f}")
    _LAST_MODEL["weights"], _LAST_MODEL["biases"] = weights, biases
    return None


def run_graph(model, mode: str = "train", model_path: str = "./models/mixed_snr_model"):
    """README.md:33-40 recipe: run_graph(create_model_graph()) with mode 'train' or 'testoldmodel'."""
    config = model["config"]
    training_set, test_set, training_labels, rand_labels, rand_set, _ = get_dataset(config)
    test_onehot = create_one_hot_labels(config.test_labels, config.n_classes)
    create_model_graph(config, training_set, training_labels, test_set, test_onehot, rand_set, rand_labels,
                       train_mode=(mode == "train"), model_save_path=model_path if mode == "train" else None,
                       model_load_path=model_path if mode != "train" else None)


def main() -> None:
    print("=" * 70)
    print("Mixed SNR Neural Viterbi Decoder")
    print("=" * 70)
    config = MixedSNRDecoderConfig()
    print("\nConfiguration:")
    print(f"  Block Length: {config.block_length}")
    print(f"  Number of Classes: {config.num_classes}")
    print(f"  Training SNR Range: {config.snr_db_vec[0]} to {config.snr_db_vec[-1]} dB")
    print(f"  Test SNR: {config.snr_db_vec[config.test_snr_index]} dB")
    print(f"  Training Samples: {config.num_training_samples}")
    print(f"  Test Samples: {config.num_test_samples}")
    print(f"  Use Dropout: {config.use_dropout}")
    print(f"  Network: {config.n_input} -> {config.n_hidden_1} -> {config.n_hidden_2} -> {config.n_hidden_3} -> {config.n_classes}")
    print("\nGenerating datasets...")
    training_set, test_set, training_labels, random_test_labels, random_test_set, class_bits = get_dataset(config)
    print(f"Training set shape: {training_set.shape}")
    print(f"Test set shape: {test_set.shape}")
    print("\nTraining neural network...")
    # the reference rebuilds UNSHUFFLED labels here (:908-914) although get_dataset shuffled the test
    # set; the permuted labels are used instead
    test_labels_onehot = create_one_hot_labels(config.test_labels, config.n_classes)
    create_model_graph(config=config, training_set=training_set, training_labels=training_labels, test_set=test_set,
                       test_labels=test_labels_onehot, random_test_set=random_test_set,
                       random_test_labels=random_test_labels, train_mode=True, model_save_path="./models/mixed_snr_model")
    print("\n" + "=" * 70)
    print("Decoding complete!")
    print("=" * 70)


if __name__ == "__main__":
    main()
