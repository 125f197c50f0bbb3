"""ConfigManager / Config: YAML -> nested attribute access, as production_wrapper.py:47-58,66-77,112-113 uses it."""
from __future__ import annotations

import os

DEFAULTS = {
    "model": {"block_length": 7, "rate": 2, "snr_db": 2.0, "n_hidden_1": 128, "n_hidden_2": 64, "n_hidden_3": 32},
    "training": {"learning_rate": 1e-4, "epochs": 200, "batch_size": 256, "display_step": 1, "beta_regularization": 1e-3,
                 "early_stopping_enabled": True, "early_stopping_patience": 20, "early_stopping_min_delta": 1e-5,
                 "samples_per_class_train": 500, "samples_per_class_test": 10000, "checkpoint_every": 10},
    "paths": {"logs_dir": "./logs", "checkpoints_dir": "./checkpoints", "models_dir": "./models", "results_dir": "./results"},
    "logging_level": "INFO", "logging_format": "text", "log_to_console": True, "log_to_file": True,
}


class Config:
    def __init__(self, data: dict):
        for k, v in data.items():
            setattr(self, k, Config(v) if isinstance(v, dict) else v)

    def to_dict(self) -> dict:
        return {k: (v.to_dict() if isinstance(v, Config) else v) for k, v in self.__dict__.items()}


def _merge(base: dict, over: dict) -> dict:
    out = dict(base)
    for k, v in (over or {}).items():
        out[k] = _merge(base[k], v) if isinstance(v, dict) and isinstance(base.get(k), dict) else v
    return out


class ConfigManager:
    def __init__(self, config_path: str = "config.yaml"):
        self.path = config_path
        data = {}
        if config_path and os.path.exists(config_path):
            import yaml
            with open(config_path) as f:
                data = yaml.safe_load(f) or {}
        if "logging" in data and isinstance(data["logging"], dict):     # accept a nested logging section too
            lg = data.pop("logging")
            data.update({"logging_level": lg.get("level", "INFO"), "logging_format": lg.get("format", "text"),
                         "log_to_console": lg.get("console", True), "log_to_file": lg.get("file", True)})
        self._config = Config(_merge(DEFAULTS, data))

    def get(self) -> Config:
        return self._config

    @staticmethod
    def _config_to_dict(config) -> dict:
        """What the wrapper stores next to a checkpoint (production_wrapper.py:248)."""
        return config.to_dict() if hasattr(config, "to_dict") else dict(config)
