"""MetricsTracker / EarlyStopping — production_wrapper.py:61-63,72-79,200-258,283-285.  BER / FER /
accuracy come from libnvd's fused popcount reduction (fnn_viterbi.error_counts), not from Python loops."""
from __future__ import annotations

import json
import os
from dataclasses import dataclass, asdict

import numpy as np


@dataclass
class Metrics:
    ber: float
    fer: float
    accuracy: float
    bit_errors: int
    frame_errors: int
    total_bits: int
    total_frames: int

    def to_dict(self):
        return asdict(self)

    def __str__(self):
        return f"BER={self.ber:.6f} FER={self.fer:.6f} Acc={self.accuracy:.4f}"


class MetricsTracker:
    def __init__(self, block_length: int):
        self.block_length = block_length
        self.history = []

    def calculate_metrics(self, predictions, labels, class_bits=None) -> Metrics:
        """predictions: class indices [B]; labels: class indices [B] or one-hot [B, C]."""
        import fnn_viterbi
        labels = np.asarray(labels)
        if labels.ndim == 2:
            labels = np.argmax(labels, axis=1)
        be, fe = fnn_viterbi.error_counts(labels.astype(np.int64), np.asarray(predictions).astype(np.int64), self.block_length)
        n = len(predictions)
        return Metrics(ber=be / (n * self.block_length), fer=fe / n, accuracy=1.0 - fe / n, bit_errors=be, frame_errors=fe,
                       total_bits=n * self.block_length, total_frames=n)

    def add_epoch_metrics(self, epoch: int, train_metrics=None, test_metrics=None, loss=None, epoch_time=None,
                          val_metrics=None, train_loss=None, **extra):
        """The wrapper's call (production_wrapper.py:229-234) names the test metrics `val_metrics` and the cost
        `train_loss`; both spellings are accepted."""
        test_metrics = test_metrics if test_metrics is not None else val_metrics
        loss = loss if loss is not None else train_loss
        rec = {"epoch": epoch, "loss": None if loss is None else float(loss), "epoch_time": epoch_time}
        if train_metrics is not None:
            rec["train"] = train_metrics.to_dict() if hasattr(train_metrics, "to_dict") else train_metrics
        if test_metrics is not None:
            rec["test"] = test_metrics.to_dict() if hasattr(test_metrics, "to_dict") else test_metrics
        rec.update({k: (v.to_dict() if hasattr(v, "to_dict") else v) for k, v in extra.items()})
        self.history.append(rec)

    def save_history(self, path):
        path = str(path)
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        with open(path, "w") as f:
            json.dump(self.history, f, indent=2, default=str)


class EarlyStopping:
    """Stop when the monitored value has not improved by min_delta for `patience` calls."""

    def __init__(self, patience: int = 20, min_delta: float = 0.0, mode: str = "min"):
        self.patience, self.min_delta, self.mode = patience, min_delta, mode
        self.best, self.bad, self.should_stop = None, 0, False

    def __call__(self, value: float) -> bool:
        better = self.best is None or (value < self.best - self.min_delta if self.mode == "min" else value > self.best + self.min_delta)
        if better:
            self.best, self.bad = value, 0
        else:
            self.bad += 1
            self.should_stop = self.bad >= self.patience
        return self.should_stop
