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
from utils.checkpoint import CheckpointManager, create_model_saver  # noqa: E402

from fnn_viterbi import (DecoderConfig, train_test_dataset, get_weights, multilayer_perceptron, predict,  # noqa: E402,F401
                         encode133171, modulateAwgn)
from nvd import training, weights as _weights  # noqa: E402


class ProductionTrainer:
    def __init__(self, config_path: str = "config.yaml"):
        self.config_mgr = ConfigManager(config_path)
        self.config = self.config_mgr.get()
        c = self.config
        self.logger = setup_logger(name="neural_viterbi", log_level=c.logging_level, log_dir=c.paths.logs_dir,
                                   use_json=(c.logging_format == "json"), log_to_console=c.log_to_console,
                                   log_to_file=c.log_to_file)
        self.metrics_tracker = MetricsTracker(c.model.block_length)
        self.checkpoint_mgr = CheckpointManager(checkpoint_dir=c.paths.checkpoints_dir, max_to_keep=5)
        self.checkpoint_manager = self.checkpoint_mgr
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
        """One Adam optimizer over all epochs (production_wrapper.py:141-190); after every epoch the train- and
        test-set decode + BER/FER count run on the GPU, checkpoints are written when epoch % 10 == 0 or on the last
        epoch (:237-250), and early stopping watches the test BER (:253-258)."""
        c = self.config.training
        model = _weights.from_dicts(self.weights, self.biases, ["tanh"] * 4)
        saver = create_model_saver(max_to_keep=5)
        state = {"t0": time.time(), "last": -1}

        def after_epoch(epoch, cost, m):
            self.weights, self.biases = _weights.to_dicts(m)
            saver.weights, saver.biases = self.weights, self.biases
            state["last"] = epoch
            if epoch % max(c.display_step, 1) != 0:
                state["t0"] = time.time()
                return False
            train_m = self._evaluate(self.train_x, self.train_y)
            test_m = self._evaluate(self.test_x, self.test_y)
            dt = time.time() - state["t0"]
            self.logger.info(f"Epoch {epoch + 1}/{c.epochs} | Time: {dt:.2f}s | Loss: {cost:.6f}")
            self.logger.info(f"  Train: {train_m}")
            self.logger.info(f"  Test:  {test_m}")
            self.metrics_tracker.add_epoch_metrics(epoch=epoch + 1, train_metrics=train_m, val_metrics=test_m, train_loss=cost,
                                                   epoch_time=dt)
            if epoch % 10 == 0 or epoch == c.epochs - 1:
                path = self.checkpoint_mgr.save_checkpoint(
                    saver=saver, session=None, epoch=epoch + 1,
                    metrics={"ber": test_m.ber, "fer": test_m.fer, "accuracy": test_m.accuracy, "loss": cost},
                    config=self.config_mgr._config_to_dict(self.config))
                self.logger.info(f"Checkpoint saved: {path}")
            state["t0"] = time.time()
            if self.early_stopping is not None and self.early_stopping(test_m.ber):
                self.logger.info(f"Early stopping triggered at epoch {epoch + 1}")
                return True
            return False

        training.train_mlp(model, self.train_x, self.train_y, acts=model.acts, loss="softmax_ce", beta=c.beta_regularization,
                           l2_layers=(0, 1, 2), lr=c.learning_rate, epochs=c.epochs, batch_size=c.batch_size, display_step=0,
                           epoch_callback=after_epoch)
        self.weights, self.biases = _weights.to_dicts(model)
        self.logger.info("Training completed!")
        final = self._evaluate(self.test_x, self.test_y)
        self.logger.info(f"Final metrics: {final}")
        os.makedirs(self.config.paths.models_dir, exist_ok=True)
        final_path = os.path.join(self.config.paths.models_dir, "final_model.npz")
        _weights.save_mlp(final_path, model)
        self.logger.info(f"Final model saved: {final_path}")
        history_path = os.path.join(self.config.paths.results_dir, "training_history.json")
        self.metrics_tracker.save_history(history_path)
        self.logger.info(f"Metrics history saved: {history_path}")
        best = self.checkpoint_mgr.get_best_checkpoint(metric="ber")
        if best:
            self.logger.info(f"Best checkpoint: Epoch {best['epoch']} with BER {best['metrics']['ber']:.6f}")
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
