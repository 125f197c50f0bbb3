#!/usr/bin/env python
"""Drop-in for /root/reference/production_wrapper.py: `ProductionTrainer(config_path)` with chainable
prepare_data() / build_model() / train() and run() -> 0 | 1, YAML config, logger, per-epoch BER/FER
metrics, periodic checkpoints, early stopping on the test BER and a history JSON (:41-307).

Training is torch autograd (off the hot path); every forward pass that produces a reported metric
(train-set / test-set decode each epoch, final evaluation) and the BER/FER counting run in libnvd's
CUDA kernels.  The reference imports a `utils` package and three camelCase names that do not
exist in its tree (:26-38); they are supplied here (utils/, fnn_viterbi aliases)."""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
for _p in (_HERE, os.path.join(_HERE, "src")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

from utils.config import ConfigManager  # noqa: E402
from utils.logger import setup_logger  # noqa: E402
from utils.metrics import MetricsTracker, EarlyStopping  # noqa: E402
from utils.checkpoint import CheckpointManager  # noqa: E402

from fnn_viterbi import (DecoderConfig, train_test_dataset, get_weights, multilayer_perceptron, predict,  # noqa: E402,F401
                         encode133171, modulateAwgn)
from nvd import training, weights as _weights  # noqa: E402


class ProductionTrainer:
    def __init__(self, config_path: str = "config.yaml"):
        self.config = ConfigManager(config_path).get()
        c = self.config
        self.logger = setup_logger(name="neural_viterbi", log_level=c.logging_level, log_dir=c.paths.logs_dir,
                                   use_json=(c.logging_format == "json"), log_to_console=c.log_to_console,
                                   log_to_file=c.log_to_file)
        self.metrics_tracker = MetricsTracker(c.model.block_length)
        self.checkpoint_manager = CheckpointManager(c.paths.checkpoints_dir)
        self.early_stopping = (EarlyStopping(patience=c.training.early_stopping_patience,
                                             min_delta=c.training.early_stopping_min_delta, mode="min")
                               if c.training.early_stopping_enabled else None)
        self.decoder_config = DecoderConfig(block_length=c.model.block_length, rate=c.model.rate, snr_db=c.model.snr_db,
                                            n_hidden_1=c.model.n_hidden_1, n_hidden_2=c.model.n_hidden_2,
                                            n_hidden_3=c.model.n_hidden_3, learning_rate=c.training.learning_rate,
                                            training_epochs=c.training.epochs, batch_size=c.training.batch_size,
                                            beta_regularization=c.training.beta_regularization,
                                            samples_per_class_train=c.training.samples_per_class_train,
                                            samples_per_class_test=c.training.samples_per_class_test)
        self.weights = self.biases = None
        self.logger.info("ProductionTrainer initialised")

    def prepare_data(self):
        self.logger.info("Preparing datasets (encoder + channel on the GPU)...")
        (self.train_x, self.train_y, self.test_x, self.test_y, self.class_bits) = train_test_dataset(self.decoder_config)
        self.logger.info(f"train {self.train_x.shape}, test {self.test_x.shape}")
        return self

    def build_model(self):
        self.weights, self.biases = get_weights(self.decoder_config)
        self.logger.info("Model built: " + " -> ".join(str(d) for d in
                                                       [self.decoder_config.n_input, self.decoder_config.n_hidden_1,
                                                        self.decoder_config.n_hidden_2, self.decoder_config.n_hidden_3,
                                                        self.decoder_config.n_classes]))
        return self

    def _evaluate(self, x, y_onehot):
        pred = predict(x, self.weights, self.biases)
        return self.metrics_tracker.calculate_metrics(pred, y_onehot, self.class_bits)

    def train(self):
        c = self.config.training
        model = _weights.from_dicts(self.weights, self.biases, ["tanh"] * 4)
        every = getattr(c, "checkpoint_every", 10)
        for epoch in range(c.epochs):
            t0 = time.time()
            cost = training.train_mlp(model, self.train_x, self.train_y, acts=model.acts, loss="softmax_ce",
                                      beta=c.beta_regularization, l2_layers=(0, 1, 2), lr=c.learning_rate, epochs=1,
                                      batch_size=c.batch_size, display_step=0)[0]
            self.weights, self.biases = _weights.to_dicts(model)
            train_m = self._evaluate(self.train_x, self.train_y)
            test_m = self._evaluate(self.test_x, self.test_y)
            dt = time.time() - t0
            self.metrics_tracker.add_epoch_metrics(epoch + 1, train_m, test_m, loss=cost, epoch_time=dt)
            if epoch % max(c.display_step, 1) == 0:
                self.logger.info(f"Epoch {epoch + 1:04d} | cost {cost:.6f} | train BER {train_m.ber:.6f} | "
                                 f"test BER {test_m.ber:.6f} FER {test_m.fer:.6f} | {dt:.2f}s")
            if (epoch + 1) % every == 0:
                self.checkpoint_manager.save_checkpoint((self.weights, self.biases), None, epoch + 1, test_m, self.config)
            if self.early_stopping is not None and self.early_stopping(test_m.ber):
                self.logger.info(f"Early stopping at epoch {epoch + 1}")
                break
        final = self._evaluate(self.test_x, self.test_y)
        self.logger.info(f"Final test BER {final.ber:.6f} FER {final.fer:.6f} accuracy {final.accuracy:.4f}")
        os.makedirs(self.config.paths.models_dir, exist_ok=True)
        _weights.save_mlp(os.path.join(self.config.paths.models_dir, "final_model.npz"), model)
        self.metrics_tracker.save_history(os.path.join(self.config.paths.results_dir, "training_history.json"))
        best = self.checkpoint_manager.get_best_checkpoint(metric="ber")
        if best:
            self.logger.info(f"Best checkpoint: {best}")
        self.final_metrics = final
        return self

    def run(self) -> int:
        try:
            self.prepare_data().build_model().train()
            return 0
        except Exception as e:  # same blanket behaviour as production_wrapper.py:299-307
            self.logger.error(f"Training failed: {e}", exc_info=True)
            return 1


def main() -> int:
    ap = argparse.ArgumentParser(description="Train the neural Viterbi decoder (production wrapper)")
    ap.add_argument("--config", default="config.yaml")
    args = ap.parse_args()
    return ProductionTrainer(args.config).run()


if __name__ == "__main__":
    sys.exit(main())
