"""Drop-in for /root/reference/src/fnn_viterbi_seq_predict.py: sliding-window decoder over one long
stream.  A window of 200 noisy symbols (stride 2) -> MLP 200 -> 500 x4 -> 2 (tanh on all five
layers) -> argmax = the info bit whose two symbols open the window (:325-334).  Encoder, channel and
the decoder forward run in libnvd (fp32 layer kernel, or the wide-layer tcgen05 kernel in bf16 mode; SURVEY.md §8f-2 — this net is
not part of the tensor-core fused path yet)."""
from __future__ import annotations

import os
from typing import Dict, Optional, Tuple

import numpy as np

from nvd import _lib, runtime, training, weights as _weights


class SequentialDecoderConfig:
    """Attribute bag identical to /root/reference/src/fnn_viterbi_seq_predict.py:35-126."""

    def __init__(self, total_train_length_info: int = 1000100, total_test_length_info: int = 1000100,
                 window_length: int = 200, n_hidden_1: int = 500, n_hidden_2: int = 500, n_hidden_3: int = 500,
                 n_hidden_4: int = 500, learning_rate: float = 0.0004, training_epochs: int = 50, batch_size: int = 100,
                 beta_regularization: float = 0.0, snr_db_range: Tuple[int, int] = (-2, 9), test_snr_index: int = 1,
                 model_save_path: str = "./models/seq_predict_model", model_load_path: Optional[str] = None,
                 test_old_model: bool = False, quantize: str = "trunc_int32"):
        self.total_train_length_info = total_train_length_info
        self.total_train_length_coded = 2 * total_train_length_info
        self.total_test_length_info = total_test_length_info
        self.total_test_length_coded = 2 * total_test_length_info
        self.window_length = window_length
        self.training_size = 1
        self.test_size = 1
        self.total_training_instances = int((self.total_train_length_coded - window_length) * 0.5)
        self.total_test_instances = int((self.total_test_length_coded - window_length) * 0.5)
        self.n_input = window_length
        self.n_hidden_1, self.n_hidden_2, self.n_hidden_3, self.n_hidden_4 = n_hidden_1, n_hidden_2, n_hidden_3, n_hidden_4
        self.n_output = 2
        self.learning_rate = learning_rate
        self.training_epochs = training_epochs
        self.batch_size = batch_size
        self.beta_regularization = beta_regularization
        self.display_step = 1
        snr_db_vec = np.arange(snr_db_range[0], snr_db_range[1])
        self.snr_db_vec = snr_db_vec
        self.snr_vec = 10 ** (snr_db_vec / 10.0)
        self.n0_vec = 1.0 / self.snr_vec
        self.sigma_vec = np.sqrt(self.n0_vec * 0.5)
        self.test_snr_index = test_snr_index
        self.model_save_path = model_save_path
        self.model_load_path = model_load_path
        self.test_old_model = test_old_model
        self.quantize = quantize     # int32 destination arrays truncate the channel output (:243-250, B-4)


def encode_133171(bits: np.ndarray) -> np.ndarray:
    return runtime.get_sim().encode_bits(np.asarray(bits), "133171")


def modulate_awgn(encoded_seq: np.ndarray, sigma: float) -> np.ndarray:
    enc = np.asarray(encoded_seq)
    rows = enc.reshape(-1, enc.shape[-1])
    f0 = runtime.take_frames(rows.shape[0])
    return runtime.get_sim().channel_bits(rows != 0, float(sigma), 0, f0, "f32").astype(np.float64).reshape(enc.shape)


def _stream(total_length_info: int, size: int, sigmas, quantize: str = "trunc_int32"):
    sim = runtime.get_sim()
    info = np.random.randint(2, size=(size, total_length_info))
    coded = sim.encode_bits(info, "133171")
    out = np.empty(coded.shape, dtype=np.int64 if quantize == "trunc_int32" else np.float32)
    for i in range(size):
        f0 = runtime.take_frames(1)
        out[i] = sim.channel_bits(coded[i], float(sigmas[i]), 0, f0, "i32_trunc" if quantize == "trunc_int32" else "f32")
    return info, out


def generate_training_data(total_length_info: int, training_size: int, sigma_vec: np.ndarray):
    """One long stream per row, a random SNR per stream (:216-252)."""
    sig = [sigma_vec[np.random.randint(len(sigma_vec))] for _ in range(training_size)]
    info, enc = _stream(total_length_info, training_size, sig)
    return info.astype(np.float32), enc


def generate_test_data(total_length_info: int, test_size: int, sigma_vec: np.ndarray, noise_index: int):
    return _stream(total_length_info, test_size, [sigma_vec[noise_index]] * test_size)


def get_single_instance(window_length: int, coded_seq: np.ndarray, info_seq: np.ndarray, offset: int):
    """Window coded[offset : offset + window_length] and the one-hot label of info bit offset / 2 (:293-336)."""
    coded_t, info_t = np.transpose(coded_seq), np.transpose(info_seq)
    batch_x = coded_t[offset:offset + window_length, 0]
    batch_y = np.zeros(2, dtype=np.float32)
    batch_y[int(info_t[int(offset * 0.5), 0])] = 1.0
    return batch_x, batch_y


def windows(coded_row: np.ndarray, window_length: int, count: int, start: int = 0) -> np.ndarray:
    """`count` stride-2 windows starting at window index `start` as a [count, window_length] array (im2col)."""
    idx = 2 * (start + np.arange(count))[:, None] + np.arange(window_length)[None, :]
    return np.asarray(coded_row)[idx]


def build_network(x, config: SequentialDecoderConfig, weights: Optional[Dict] = None, biases: Optional[Dict] = None):
    """-> (output [B, 2], weights, biases): tanh after all five affines (:408-412), evaluated on the GPU."""
    if weights is None:
        dims = [config.n_input, config.n_hidden_1, config.n_hidden_2, config.n_hidden_3, config.n_hidden_4, config.n_output]
        weights, biases = _weights.to_dicts(_weights.init_mlp(dims, ["tanh"] * 5))
    out = None
    if x is not None:
        _, out = _forward(x, weights, biases)
    return out, weights, biases


def _forward(x, weights, biases):
    sim = runtime.get_sim()
    h = runtime.cached_mlp(_weights.from_dicts(weights, biases, ["tanh"] * 5))
    # fp32 parity path by default; runtime.set_precision("bf16") / NVD_PRECISION=bf16 selects the wide-layer
    # tcgen05 kernel (18x faster on B200, 2e-2 tolerance)
    prec = "bf16" if runtime.precision() == "bf16" and sim.mlp_tc_supported(h) else "fp32"
    return sim.decode(h, np.asarray(x, dtype=np.float32), prec)


def train_and_evaluate(config: SequentialDecoderConfig, train_info, train_encoded, test_info, test_encoded,
                       weights: Optional[Dict] = None, biases: Optional[Dict] = None, chunk: int = 1 << 16):
    """Train over the stride-2 windows of the training stream (:517-580), then decode every test window
    on the GPU and count bit errors (:590-616).  -> (predictions, error_rate)."""
    if config.test_old_model and config.model_load_path:
        weights, biases = _weights.to_dicts(_weights.load_mlp(config.model_load_path))
    elif weights is None:
        _, weights, biases = build_network(None, config)
    if not config.test_old_model and config.training_epochs > 0:
        model = _weights.from_dicts(weights, biases, ["tanh"] * 5)
        n = config.total_training_instances
        for epoch in range(config.training_epochs):
            cost, done = 0.0, 0
            for s in range(0, n, chunk):
                c = min(chunk, n - s)
                x = windows(train_encoded[0], config.window_length, c, s)
                y = np.eye(2, dtype=np.float32)[np.asarray(train_info[0][s:s + c]).astype(int)]
                cs = training.train_mlp(model, x, y, acts=model.acts, loss="softmax_ce", beta=config.beta_regularization,
                                        l2_layers=(0, 1, 2, 3), lr=config.learning_rate, epochs=1,
                                        batch_size=config.batch_size, display_step=0)
                cost += cs[0] * c
                done += c
            print(f"Epoch: {epoch + 1:04d} | cost = {cost / max(done, 1):.9f}")
        print("Optimization Finished!")
        weights, biases = _weights.to_dicts(model)
        if config.model_save_path:
            os.makedirs(os.path.dirname(os.path.abspath(config.model_save_path)), exist_ok=True)
            _weights.save_mlp(config.model_save_path + ".npz", model)
    n = config.total_test_instances
    predictions = np.empty(n, dtype=np.int64)
    for s in range(0, n, chunk):
        c = min(chunk, n - s)
        pred, _ = _forward(windows(test_encoded[0], config.window_length, c, s), weights, biases)
        predictions[s:s + c] = pred
    errors = int(np.sum(predictions != np.asarray(test_info[0][:n]).astype(int)))
    error_rate = errors / n
    print(f"Test Error Rate: {error_rate:.6f}")
    print(f"Test Accuracy: {1 - error_rate:.6f}")
    return predictions, error_rate


def main() -> None:
    print("=" * 70)
    print("Sequential (sliding-window) Neural Viterbi Decoder")
    print("=" * 70)
    config = SequentialDecoderConfig()
    train_info, train_encoded = generate_training_data(config.total_train_length_info, config.training_size, config.sigma_vec)
    test_info, test_encoded = generate_test_data(config.total_test_length_info, config.test_size, config.sigma_vec,
                                                 config.test_snr_index)
    train_and_evaluate(config, train_info, train_encoded, test_info, test_encoded)
    print("Decoding complete!")


if __name__ == "__main__":
    main()
