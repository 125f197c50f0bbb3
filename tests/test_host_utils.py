"""CPU tests of the host-side pieces either side of the path: the `utils` package as production_wrapper.py calls it
(/root/reference/production_wrapper.py:47-79,200-250,283-293), the training loop's single optimizer, the model-cache key."""
import json
import os

import numpy as np
import pytest


def test_utils_api_matches_the_wrappers_call_sites(tmp_path):
    from utils.checkpoint import CheckpointManager, create_model_saver
    from utils.config import ConfigManager
    from utils.metrics import EarlyStopping, Metrics, MetricsTracker
    from nvd import weights
    mgr = ConfigManager(str(tmp_path / "missing.yaml"))
    cfg = mgr.get()
    assert cfg.model.block_length == 7 and cfg.training.early_stopping_enabled in (True, False)
    d = mgr._config_to_dict(cfg)                                   # production_wrapper.py:248
    assert d["model"]["block_length"] == 7 and json.dumps(d)
    saver = create_model_saver(max_to_keep=5)                       # :147
    m = weights.init_mlp([14, 128, 64, 32, 128], ["tanh"] * 4, seed=1)
    saver.weights, saver.biases = weights.to_dicts(m)
    ck = CheckpointManager(checkpoint_dir=str(tmp_path / "ck"), max_to_keep=5)     # :66-69
    for epoch, ber in ((1, 0.2), (11, 0.05), (21, 0.07)):
        path = ck.save_checkpoint(saver=saver, session=None, epoch=epoch,
                                  metrics={"ber": ber, "fer": 2 * ber, "accuracy": 1 - 2 * ber, "loss": np.float32(1.5)}, config=d)
        assert os.path.exists(path)
    best = ck.get_best_checkpoint(metric="ber")                     # :288-292 indexes the result
    assert best["epoch"] == 11 and best["metrics"]["ber"] == pytest.approx(0.05)
    w, b = ck.load(best["path"])
    assert np.array_equal(w["h1"], saver.weights["h1"])
    final = saver.save(None, str(tmp_path / "final_model"))         # :278-280
    assert os.path.exists(final)
    tr = MetricsTracker(block_length=7)                             # :61-63
    mt = Metrics(ber=0.1, fer=0.2, accuracy=0.8, bit_errors=7, frame_errors=2, total_bits=70, total_frames=10)
    tr.add_epoch_metrics(epoch=1, train_metrics=mt, val_metrics=mt, train_loss=np.float64(0.5))    # :229-234
    tr.save_history(tmp_path / "res" / "history.json")              # a pathlib.Path, as the wrapper passes (:283-285)
    hist = json.load(open(tmp_path / "res" / "history.json"))
    assert hist[0]["test"]["ber"] == 0.1 and hist[0]["loss"] == 0.5 and "BER=" in str(mt)
    es = EarlyStopping(patience=2, min_delta=0.0, mode="min")       # :72-79
    assert [es(v) for v in (0.5, 0.4, 0.45, 0.41)] == [False, False, False, True]


def test_training_keeps_one_optimizer_across_epochs():
    """train_mlp(epochs=N, epoch_callback=...) is ONE Adam run (the reference keeps a single AdamOptimizer alive,
    fnn_viterbi.py:562-603): it equals the plain N-epoch call and differs from N fresh one-epoch calls."""
    torch = pytest.importorskip("torch")
    from nvd import training, weights
    rng = np.random.default_rng(0)
    x = rng.standard_normal((512, 14)).astype(np.float32)
    y = np.eye(128, dtype=np.float32)[rng.integers(0, 128, 512)]
    def fresh():
        return weights.init_mlp([14, 128, 64, 32, 128], ["tanh"] * 4, seed=5)
    kw = dict(acts=[1, 1, 1, 1], loss="softmax_ce", beta=1e-3, lr=1e-3, batch_size=64, display_step=0, backend="torch")
    a, b, c = fresh(), fresh(), fresh()
    torch.manual_seed(0)
    costs_a = training.train_mlp(a, x, y, epochs=3, **kw)
    seen = []
    costs_b = training.train_mlp(b, x, y, epochs=3, epoch_callback=lambda e, cost, m: seen.append((e, cost, m.Ws[0].copy())) and False, **kw)
    assert np.allclose(costs_a, costs_b) and all(np.array_equal(u, v) for u, v in zip(a.Ws, b.Ws))
    assert [e for e, _, _ in seen] == [0, 1, 2] and np.array_equal(seen[-1][2], b.Ws[0])
    for _ in range(3):
        training.train_mlp(c, x, y, epochs=1, **kw)                 # Adam moments reset every call
    assert not np.allclose(c.Ws[0], a.Ws[0])
    stop = fresh()
    assert len(training.train_mlp(stop, x, y, epochs=5, epoch_callback=lambda e, cost, m: e == 1, **kw)) == 2   # early stop


def test_model_cache_key_is_the_parameter_content():
    from nvd import runtime, weights
    m1 = weights.init_mlp([14, 20, 4], ["tanh", "tanh"], seed=1)
    m2 = weights.init_mlp([14, 20, 4], ["tanh", "tanh"], seed=1)
    k = runtime._digest(list(m1.Ws) + list(m1.bs), m1.acts)
    assert k == runtime._digest(list(m2.Ws) + list(m2.bs), m2.acts)
    m2.Ws[1][0, 0] += 1e-3
    assert k != runtime._digest(list(m2.Ws) + list(m2.bs), m2.acts)
    assert k != runtime._digest(list(m1.Ws) + list(m1.bs), [1, 2])
