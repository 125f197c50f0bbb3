"""CheckpointManager — production_wrapper.py:66-69,237-250,288-293.  The reference hands a
tf.train.Saver + session; here `saver` is any object with `.weights` / `.biases` dicts (or a
(weights, biases) tuple) and checkpoints are `.npz` files keyed by the reference's TF variable names."""
from __future__ import annotations

import json
import os

import numpy as np


class ModelSaver:
    """Stand-in for the tf.train.Saver the reference creates (production_wrapper.py:147): holds the current
    parameter dicts; `save(session, path)` writes them as `.npz` keyed by the reference's TF variable names."""

    def __init__(self, max_to_keep: int = 5):
        self.max_to_keep = max_to_keep
        self.weights, self.biases = None, None

    def save(self, session=None, path: str = "model") -> str:
        from nvd import weights as W
        path = str(path)
        path = path if path.endswith(".npz") else path + ".npz"
        W.save_mlp(path, W.from_dicts(self.weights, self.biases, ["tanh"] * len(self.weights)))
        return path


def create_model_saver(max_to_keep: int = 5) -> ModelSaver:
    return ModelSaver(max_to_keep)


class CheckpointManager:
    def __init__(self, checkpoint_dir: str = "./checkpoints", max_to_keep: int = 5):
        self.dir, self.max_to_keep = checkpoint_dir, max_to_keep
        os.makedirs(checkpoint_dir, exist_ok=True)
        self.index_path = os.path.join(checkpoint_dir, "index.json")
        self.index = json.load(open(self.index_path)) if os.path.exists(self.index_path) else []

    def save_checkpoint(self, saver, session=None, epoch: int = 0, metrics=None, config=None) -> str:
        from nvd import weights as W
        weights, biases = (saver.weights, saver.biases) if hasattr(saver, "weights") else saver
        path = os.path.join(self.dir, f"ckpt_epoch_{epoch:04d}.npz")
        W.save_mlp(path, W.from_dicts(weights, biases, ["tanh"] * len(weights)))
        m = metrics.to_dict() if hasattr(metrics, "to_dict") else (metrics or {})
        entry = {"epoch": epoch, "path": path, "metrics": {k: (float(v) if isinstance(v, (int, float, np.floating, np.integer)) else v)
                                                             for k, v in dict(m).items()}}
        if config is not None:
            entry["config"] = config if isinstance(config, dict) else (config.to_dict() if hasattr(config, "to_dict") else None)
        self.index.append(entry)
        while len(self.index) > self.max_to_keep:
            old = self.index.pop(0)
            if os.path.exists(old["path"]):
                os.remove(old["path"])
        with open(self.index_path, "w") as f:
            json.dump(self.index, f, indent=2, default=str)
        return path

    def get_best_checkpoint(self, metric: str = "ber", mode: str = "min"):
        scored = [e for e in self.index if metric in e.get("metrics", {})]
        if not scored:
            return None
        # the whole index entry, as the wrapper reads it: best['epoch'], best['metrics']['ber'] (:288-292), best['path']
        return (min if mode == "min" else max)(scored, key=lambda e: e["metrics"][metric])

    def load(self, path: str):
        from nvd import weights as W
        return W.to_dicts(W.load_mlp(path))
