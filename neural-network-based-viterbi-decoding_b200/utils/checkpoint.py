"""CheckpointManager — production_wrapper.py:66-69,237-250,288-293.  The reference hands a
tf.train.Saver + session; here `saver` is any object with `.weights` / `.biases` dicts (or a
(weights, biases) tuple) and checkpoints are `.npz` files keyed by the reference's TF variable names."""
from __future__ import annotations

import json
import os

import numpy as np


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
        self.index.append({"epoch": epoch, "path": path, "metrics": m})
        while len(self.index) > self.max_to_keep:
            old = self.index.pop(0)
            if os.path.exists(old["path"]):
                os.remove(old["path"])
        with open(self.index_path, "w") as f:
            json.dump(self.index, f, indent=2)
        return path

    def get_best_checkpoint(self, metric: str = "ber", mode: str = "min"):
        scored = [e for e in self.index if metric in e.get("metrics", {})]
        if not scored:
            return None
        best = (min if mode == "min" else max)(scored, key=lambda e: e["metrics"][metric])
        return best["path"]

    def load(self, path: str):
        from nvd import weights as W
        return W.to_dicts(W.load_mlp(path))
