"""Drop-in for /root/reference/src/fnn_viterbi_bitwise.py: per-bit predictor.  6 state bits + 512 info
bits -> 1036 noisy symbols -> MLP 1036 -> 30 (ReLU) -> 2 (ReLU) -> argmax = value of info bit
`bit_position`.  Encoder, channel, forward pass, argmax and the error count run in libnvd's CUDA
kernels; only the wired graph is built (h2-h4 are created by the reference but unused, B-11)."""
from __future__ import annotations

import os
from typing import Dict, Optional, Tuple

import numpy as np

from nvd import _lib, runtime, training, weights as _weights


class BitwiseDecoderConfig:
    """Attribute bag identical to /root/reference/src/fnn_viterbi_bitwise.py:29-107 (+ `quantize`)."""

    def __init__(self, constraint_length: int = 6, block_length: int = 512, n_hidden_1: int = 30, n_hidden_2: int = 100,
                 n_hidden_3: int = 100, n_hidden_4: int = 100, learning_rate: float = 0.0002, training_epochs: int = 300,
                 batch_size: int = 1000, beta_regularization: float = 0.0001, training_size: int = 300000,
                 test_size: int = 300000, snr_db_range: Tuple[int, int] = (-2, 9), test_snr_index: int = 0,
                 bit_position: int = 0, starting_state: int = 0, quantize: str = "trunc_int32"):
        self.constraint_length = constraint_length
        self.block_length = block_length
        self.total_length = constraint_length + block_length
        self.n_input = 2 * self.total_length
        self.n_hidden_1, self.n_hidden_2, self.n_hidden_3, self.n_hidden_4 = n_hidden_1, n_hidden_2, n_hidden_3, n_hidden_4
        self.n_output = 2
        self.learning_rate = learning_rate
        self.training_epochs = training_epochs
        self.batch_size = batch_size
        self.beta_regularization = beta_regularization
        self.display_step = 1
        self.training_size = training_size
        self.test_size = test_size
        snr_db_vec = np.arange(snr_db_range[0], snr_db_range[1])
        self.snr_db_vec = snr_db_vec
        self.snr_vec = 10 ** (snr_db_vec / 10.0)
        self.n0_vec = 1.0 / self.snr_vec
        self.sigma_vec = np.sqrt(self.n0_vec * 0.5)
        self.test_snr_index = test_snr_index
        self.bit_position = bit_position
        self.starting_state = starting_state
        # 'trunc_int32': the reference stores channel outputs into an int32 array, i.e. truncates them
        # toward zero (:274-285, :309-317); 'none' keeps float32.
        self.quantize = quantize


def extract_state_bits(state_number: int, constraint_length: int = 6) -> np.ndarray:
    """MSB-first binary of the state number; states beyond 2**constraint_length are masked (B-6)."""
    s = int(state_number) & ((1 << constraint_length) - 1)
    return np.array([(s >> (constraint_length - 1 - i)) & 1 for i in range(constraint_length)], dtype=np.int32)


def encode_133171_with_state(bits: np.ndarray, state_num: int) -> np.ndarray:
    """State bits prepended, the whole (L+6) sequence encoded from zero state on the GPU (:162-163)."""
    b = np.mod(np.asarray(bits), 2).astype(np.int64)
    return runtime.get_sim().encode_bits(np.concatenate([extract_state_bits(state_num, 6), b]), "133171")


def modulate_awgn(encoded_seq: np.ndarray, sigma: float) -> np.ndarray:
    enc = np.asarray(encoded_seq)
    rows = enc.reshape(-1, enc.shape[-1])
    f0 = runtime.take_frames(rows.shape[0])
    return runtime.get_sim().channel_bits(rows != 0, float(sigma), 0, f0, "f32").astype(np.float64).reshape(enc.shape)


def generate_input_vectors(length: int, number: int) -> np.ndarray:
    return np.random.randint(2, size=(number, length))


def _encode_and_transmit(config: BitwiseDecoderConfig, info: np.ndarray, sigma_per_frame: np.ndarray) -> np.ndarray:
    sim = runtime.get_sim()
    B = info.shape[0]
    pre = np.tile(extract_state_bits(config.starting_state, config.constraint_length), (B, 1))
    coded = sim.encode_bits(np.concatenate([pre, info.astype(np.int64)], axis=1), "133171")      # [B, 2(L+6)]
    dtype = "i32_trunc" if config.quantize == "trunc_int32" else "f32"
    out = np.empty(coded.shape, dtype=np.int64 if dtype == "i32_trunc" else np.float32)
    for sigma in np.unique(sigma_per_frame):
        idx = np.nonzero(sigma_per_frame == sigma)[0]
        f0 = runtime.take_frames(len(idx))
        out[idx] = sim.channel_bits(coded[idx], float(sigma), 0, f0, dtype)
    return out


def generate_training_data(config: BitwiseDecoderConfig) -> Tuple[np.ndarray, np.ndarray]:
    """Random SNR per frame from sigma_vec (:280-285)."""
    info = generate_input_vectors(config.block_length, config.training_size).astype(np.float32)
    sig = config.sigma_vec[np.random.randint(len(config.sigma_vec), size=config.training_size)]
    return info, _encode_and_transmit(config, info, sig)


def generate_test_data(config: BitwiseDecoderConfig) -> Tuple[np.ndarray, np.ndarray]:
    info = generate_input_vectors(config.block_length, config.test_size)
    sig = np.full(config.test_size, config.sigma_vec[config.test_snr_index])
    return info, _encode_and_transmit(config, info, sig)


def build_network(x, config: BitwiseDecoderConfig, weights: Optional[Dict] = None, biases: Optional[Dict] = None):
    """-> (output [B, 2], weights, biases): relu(relu(x W1 + b1) Wout + bout) evaluated on the GPU.  The
    dicts carry the reference's keys incl. the unused h2..h4 / b2..b4 (:347-384)."""
    if weights is None:
        rng = np.random.default_rng()
        dims = {"h1": (config.n_input, config.n_hidden_1), "h2": (config.n_hidden_1, config.n_hidden_2),
                "h3": (config.n_hidden_2, config.n_hidden_3), "h4": (config.n_hidden_3, config.n_hidden_4)}
        weights = {k: _weights.truncated_normal(s, np.sqrt(2.0 / s[0]), rng) for k, s in dims.items()}
        weights["out"] = _weights.truncated_normal((config.n_hidden_1, config.n_output), np.sqrt(2.0 / config.n_hidden_4), rng)
        biases = {"b1": np.zeros(config.n_hidden_1, np.float32), "b2": np.zeros(config.n_hidden_2, np.float32),
                  "b3": np.zeros(config.n_hidden_3, np.float32), "b4": np.zeros(config.n_hidden_4, np.float32),
                  "out": np.zeros(config.n_output, np.float32)}
    out = None
    if x is not None:
        _, out = _forward(x, weights, biases)
    return out, weights, biases


def _wired_model(weights: Dict, biases: Dict) -> _weights.MlpModel:
    return _weights.MlpModel([np.asarray(weights["h1"], np.float32), np.asarray(weights["out"], np.float32)],
                             [np.asarray(biases["b1"], np.float32), np.asarray(biases["out"], np.float32)],
                             [_lib.ACT_RELU, _lib.ACT_RELU])


def _forward(x, weights, biases):
    sim = runtime.get_sim()
    h = runtime.cached_mlp(_wired_model(weights, biases))
    # runtime.set_precision("bf16") routes to the K-streamed tcgen05 kernel (bf16 rows x bf16 W1, fp32 tail)
    return sim.decode(h, np.asarray(x, dtype=np.float32), "bf16" if runtime.precision() == "bf16" else "fp32")


def train_and_evaluate(config: BitwiseDecoderConfig, train_info, train_encoded, test_info, test_encoded,
                       model_save_path: Optional[str] = None, weights: Optional[Dict] = None, biases: Optional[Dict] = None):
    """Sigmoid-CE on the ReLU outputs + beta * (l2(h1) + l2(h2)) with Adam (:439-497), then the test-set
    forward + argmax + error count on the GPU.  -> (predictions, error_rate)."""
    if weights is None:
        _, weights, biases = build_network(None, config)
    model = _wired_model(weights, biases)
    train_labels = np.asarray(train_info)[:, config.bit_position].astype(int)
    test_labels = np.asarray(test_info)[:, config.bit_position].astype(int)
    onehot = np.zeros((len(train_labels), 2), np.float32)
    onehot[np.arange(len(train_labels)), train_labels] = 1
    reg = 0.5 * float((np.asarray(weights["h2"], np.float64) ** 2).sum())      # constant: h2 is not in the graph
    costs = training.train_mlp(model, train_encoded, onehot, acts=model.acts, loss="sigmoid_ce",
                               beta=config.beta_regularization, l2_layers=(0,), lr=config.learning_rate,
                               epochs=config.training_epochs, batch_size=config.batch_size, display_step=0)
    for epoch, c in enumerate(costs):
        if epoch % config.display_step == 0:
            print(f"Epoch: {epoch + 1:04d} | cost = {c + config.beta_regularization * reg:.9f}")
    print("Optimization Finished!")
    weights["h1"], weights["out"], biases["b1"], biases["out"] = model.Ws[0], model.Ws[1], model.bs[0], model.bs[1]
    if model_save_path:
        os.makedirs(os.path.dirname(os.path.abspath(model_save_path)), exist_ok=True)
        np.savez(model_save_path if model_save_path.endswith(".npz") else model_save_path + ".npz",
                 **{f"weights_{k}": v for k, v in weights.items()},
                 **{("bias_out" if k == "out" else "bias_h" + k[1:]): v for k, v in biases.items()})
        print(f"Model saved to {model_save_path}")
    predictions, _ = _forward(test_encoded, weights, biases)
    predictions = predictions.astype(np.int64)
    # decision against info bit `bit_position` on the GPU (K5, nvd_count_bit_errors) — :504-509
    counters = runtime.get_sim().count_bit_errors(predictions, np.asarray(test_info).astype(np.uint8), config.bit_position)
    errors = int(counters[0])
    error_rate = errors / len(predictions)
    print(f"Test Error Rate: {error_rate:.6f}")
    print(f"Test Accuracy: {1 - error_rate:.6f}")
    return predictions, error_rate


def main() -> None:
    print("=" * 70)
    print("Bitwise Neural Viterbi Decoder")
    print("=" * 70)
    config = BitwiseDecoderConfig()
    print("\nConfiguration:")
    print(f"  Block Length: {config.block_length}")
    print(f"  Constraint Length: {config.constraint_length}")
    print(f"  Predicting Bit Position: {config.bit_position}")
    print(f"  Test SNR: {config.snr_db_vec[config.test_snr_index]} dB")
    print(f"  Training Samples: {config.training_size}")
    print(f"  Test Samples: {config.test_size}")
    print("\nGenerating training data...")
    train_info, train_encoded = generate_training_data(config)
    print("Generating test data...")
    test_info, test_encoded = generate_test_data(config)
    print("\nTraining neural network...")
    train_and_evaluate(config, train_info, train_encoded, test_info, test_encoded)
    print("\n" + "=" * 70)
    print("Decoding complete!")
    print("=" * 70)


if __name__ == "__main__":
    main()
