"""Oracle: hard-decision bookkeeping and BER / FER counting (TEST INFRASTRUCTURE).

Restates
  calculate_error                    /root/reference/src/fnn_viterbi.py:619-674
  calculate_hamming_distance_decimal /root/reference/src/fnn_viterbi_mixed_snr.py:303-328
  calculate_ber / calculate_accuracy /root/reference/src/fnn_viterbi_mixed_snr.py:634-673
  binary_to_decimal (LSB-first!)     /root/reference/src/fnn_viterbi_mixed_snr.py:263-279
  extract_bits_single (MSB-first)    /root/reference/src/fnn_viterbi_mixed_snr.py:282-300
  bitwise error count                /root/reference/src/fnn_viterbi_bitwise.py:508-509
  rnn BER over the first L steps     /root/reference/src/rnn_viterbi.py:477-491
"""
from __future__ import annotations

import numpy as np


def popcount(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x).astype(np.uint64)
    n = np.zeros(x.shape, dtype=np.int64)
    while x.any():
        n += (x & np.uint64(1)).astype(np.int64)
        x = x >> np.uint64(1)
    return n


def count_errors(pred, label, L: int):
    """(bit_errors, frame_errors) for class-index decoders: popcount(pred ^ label) summed
    (== the per-class XOR loops of fnn_viterbi.py:655-667 and mixed_snr:650-655)."""
    pred = np.asarray(pred).astype(np.int64)
    label = np.asarray(label).astype(np.int64)
    x = (pred ^ label) & ((1 << L) - 1)
    return int(popcount(x).sum()), int((pred != label).sum())


def calculate_error(test_labels_onehot, predictions, class_bits, block_length,
                    samples_per_class_test=None):
    """Returns dict(accuracy, fer, ber, bit_errors, total_bits) and the four lines the
    reference prints (fnn_viterbi.py:643-644,673-674).  BER follows the reference's
    class-major assumption (:655-667): frame k belongs to class k // samples_per_class."""
    labels = np.argmax(np.asarray(test_labels_onehot), axis=1).astype(int)      # :635
    predictions = np.asarray(predictions).astype(int)
    n = len(predictions)
    frame_errors = int(np.sum(predictions != labels))                           # :639
    fer = frame_errors / n
    L = block_length
    sh = np.arange(L - 1, -1, -1)
    pred_bits = (predictions[:, None] >> sh[None, :]) & 1                        # :648-651 MSB-first
    if samples_per_class_test is None:
        true_bits = np.asarray(class_bits)[labels]
    else:
        owner = np.arange(n) // samples_per_class_test                           # :657-661
        true_bits = np.asarray(class_bits)[owner]
    bit_errors = int(np.sum(np.logical_xor(true_bits, pred_bits)))               # :667
    total_bits = n * L
    out = dict(accuracy=1.0 - fer, fer=fer, ber=bit_errors / total_bits,
               bit_errors=bit_errors, frame_errors=frame_errors, total_bits=total_bits)
    out["lines"] = [
        f"Accuracy: {out['accuracy']:.4f}",
        f"Frame Error Rate: {fer:.6f}",
        f"Bit Error Rate: {out['ber']:.6f}",
        f"Total Bit Errors: {bit_errors}/{total_bits}",
    ]
    return out


def binary_to_decimal(binary_array) -> int:
    """LSB-FIRST weights 2**arange(n) as the code does (mixed_snr:277), despite its docstring."""
    a = np.asarray(binary_array)
    return int(a.dot(2 ** np.arange(a.shape[0])))


def extract_bits_single(decimal_value: int, block_length: int) -> np.ndarray:
    """MSB-first fixed-width binary (mixed_snr:297-299)."""
    s = format(int(decimal_value), f"0{block_length}b")
    bits = np.zeros(block_length, dtype=np.uint8)
    bits[: len(s)] = [int(c) for c in s][:block_length]
    return bits


def hamming_decimal(v1: int, v2: int, block_length: int = 0) -> int:
    return bin(int(v1) ^ int(v2)).count("1")


def calculate_ber(predictions, labels, block_length: int) -> float:
    be, _ = count_errors(predictions, labels, 63)
    return be / (len(predictions) * block_length)


def calculate_accuracy(predictions, labels) -> float:
    predictions = np.asarray(predictions)
    return float(np.sum(predictions == np.asarray(labels).astype(int)) / len(predictions))


def wilson_interval(errors: int, trials: int, z: float = 1.96):
    """95 % Wilson score interval for a binomial proportion (BER/FER Monte-Carlo CI)."""
    if trials == 0:
        return 0.0, 1.0
    p = errors / trials
    d = 1 + z * z / trials
    c = p + z * z / (2 * trials)
    h = z * np.sqrt(p * (1 - p) / trials + z * z / (4 * trials * trials))
    return max(0.0, (c - h) / d), min(1.0, (c + h) / d)
