"""Drop-in module tests on the GPU: the reference's own encoder tests (tests/test_encoders.py) restated
against the CUDA-backed functions, plus data-set builders, decode and metric printing."""
import io
import contextlib

import numpy as np
import pytest

from oracle import encoders as oenc, mlp as omlp

pytestmark = pytest.mark.gpu


def test_reference_encoder_suite(golden):
    import fnn_viterbi as fv, fnn_viterbi_bitwise as fb
    rng = np.random.default_rng(0)
    for L in (7, 10, 100):                                    # test_encoders.py:44-57,68-81
        z = np.zeros(L, dtype=int)
        assert len(fv.encode_57(z)) == 2 * L and not fv.encode_57(z).any()
        assert len(fv.encode_133171(z)) == 2 * L and not fv.encode_133171(z).any()
    x = rng.integers(0, 2, 50)
    for enc in (fv.encode_57, fv.encode_133171):              # :92-106 binary-valued, deterministic
        y = enc(x)
        assert set(np.unique(y)) <= {0, 1} and y.dtype == np.int32 and (enc(x) == y).all()
    a, b = rng.integers(0, 2, 40), rng.integers(0, 2, 40)       # :169-182 GF(2) linearity
    assert ((fv.encode_133171(a) ^ fv.encode_133171(b)) == fv.encode_133171(a ^ b)).all()
    assert len(fv.encode_133171([1])) == 2 and len(fv.encode_133171([0] * 5)) == 10      # :184-209 (B-5)
    assert len(fb.encode_133171_with_state(np.ones(9, int), 3)) == 30                    # :136-142
    assert not (fb.encode_133171_with_state(x, 0) == fb.encode_133171_with_state(x, 63)).all()   # :155-163
    assert len(fb.encode_133171_with_state([1, 0, 1], 100)) == 18                       # :227-235 (B-6)
    assert (fv.encode_133171(golden["enc_float_in"]) == golden["enc_float_out"]).all()  # float inputs (% 2)
    for s in (0, 1, 33, 63):
        assert (fb.encode_133171_with_state(golden["state_in"], s) == golden["state_out"][s]).all()
    assert (fv.encode133171(x) == fv.encode_133171(x)).all()


def test_modulate_awgn_semantics():
    import fnn_viterbi as fv
    from nvd import runtime
    runtime.seed(123)
    cw = np.array([0, 1] * 7)
    y = fv.modulate_awgn(cw, 0.5)
    assert y.dtype == np.float64 and y.shape == (14,)
    runtime.seed(123)
    assert np.array_equal(fv.modulate_awgn(cw, 0.5), y)         # reproducible from (seed, frame index)
    assert not np.array_equal(fv.modulate_awgn(cw, 0.5), y)     # fresh noise on the next call
    assert np.array_equal(fv.modulate_awgn(cw, 0.0), np.array([-1.0, 1.0] * 7))
    big = fv.modulate_awgn(np.zeros((2000, 14), int), 0.7)
    assert abs(big.mean() + 1) < 0.02 and abs(big.std() - 0.7) < 0.02


def test_dataset_decode_and_error_report():
    import fnn_viterbi as fv
    from nvd import weights
    cfg = fv.DecoderConfig(samples_per_class_train=4, samples_per_class_test=50, snr_db=3.0)
    tr_x, tr_y, te_x, te_y, cb = fv.train_test_dataset(cfg)
    assert tr_x.shape == (512, 14) and te_x.shape == (6400, 14) and te_y.shape == (6400, 128) and tr_x.dtype == np.float32
    assert (np.argmax(te_y, 1) == np.arange(128).repeat(50)).all()                      # class-major
    book = 2.0 * oenc.encode_133171(cb) - 1.0
    assert np.abs(te_x - book[np.argmax(te_y, 1)]).mean() == pytest.approx(cfg.sigma * np.sqrt(2 / np.pi), rel=0.05)
    m = weights.load_default_fnn()
    w, b = weights.to_dicts(m)
    soft = fv.multilayer_perceptron(te_x, w, b)
    want = omlp.forward(te_x, m.Ws, m.bs, m.acts, dtype=np.float64)
    assert soft.shape == (6400, 128) and np.abs(soft - want).max() <= 1e-5
    pred = fv.predict(te_x, w, b)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        fv.calculate_error(te_y, pred, cb, cfg)
    lines = buf.getvalue().splitlines()
    labels = np.argmax(te_y, 1)
    fe = int((pred != labels).sum()); be = int(sum(bin(int(p) ^ int(l)).count("1") for p, l in zip(pred, labels)))
    assert lines == [f"Accuracy: {1 - fe / 6400:.4f}", f"Frame Error Rate: {fe / 6400:.6f}",
                     f"Bit Error Rate: {be / 44800:.6f}", f"Total Bit Errors: {be}/44800"]
    assert fe / 6400 < 0.05                                                             # trained decoder at 3 dB
    pred16 = fv.predict(te_x, w, b, precision="bf16")                                   # tcgen05 path
    assert (pred16 != pred).mean() < 0.01


def test_run_network_trains_and_decodes(capsys):
    import fnn_viterbi as fv
    cfg = fv.DecoderConfig(training_epochs=2, samples_per_class_train=20, samples_per_class_test=10, learning_rate=1e-3)
    tr_x, tr_y, te_x, te_y, cb = fv.train_test_dataset(cfg)
    pred = fv.run_network(cfg, tr_x, tr_y, te_x, te_y)
    out = capsys.readouterr().out
    assert "Epoch: 0001 | cost = " in out and "Optimization Finished!" in out and "Accuracy: " in out
    assert pred.shape == (1280,) and pred.min() >= 0 and pred.max() < 128


def test_mixed_snr_eval_branch(tmp_path, capsys):
    import fnn_viterbi_mixed_snr as fm
    from nvd import weights
    cfg = fm.MixedSNRDecoderConfig(samples_per_class_train=24, samples_per_class_test=20, test_snr_index=6,
                                   random_test_set_size=50)
    tr, te, tr_y, rnd_y, rnd_x, cb = fm.get_dataset(cfg)
    assert tr.shape == (3072, 14) and te.shape == (2560, 14) and rnd_x.shape == (50, 14) and rnd_y.shape == (50, 128)
    assert cfg.samples_per_snr == 2 and (tr.reshape(128, 24, 14)[:, :24] != 0).any()
    path = str(tmp_path / "model")
    w, b = weights.to_dicts(weights.load_default_fnn())
    fm.save_model(path, w, b)
    te_y = fm.create_one_hot_labels(cfg.test_labels, 128)
    fm.create_model_graph(cfg, tr, tr_y, te, te_y, rnd_x, rnd_y, train_mode=False, model_load_path=path)
    out = capsys.readouterr().out
    assert f"Model loaded from {path}" in out and "Test Set Accuracy: " in out and "Test Set BER: " in out
    acc = float(out.split("Test Set Accuracy: ")[1].split()[0])
    assert acc > 0.97                                          # 4 dB with the trained fixture
    pred = fm.predict(te, w, b)
    assert fm.calculate_ber(pred, cfg.test_labels, 7) == pytest.approx(
        sum(bin(int(p) ^ int(l)).count("1") for p, l in zip(pred, cfg.test_labels)) / (2560 * 7))


def test_bitwise_pipeline(capsys):
    import fnn_viterbi_bitwise as fb
    cfg = fb.BitwiseDecoderConfig(block_length=64, training_size=400, test_size=300, training_epochs=2, batch_size=100)
    tr_i, tr_x = fb.generate_training_data(cfg)
    te_i, te_x = fb.generate_test_data(cfg)
    assert tr_x.shape == (400, 140) and te_x.shape == (300, 140)
    assert np.array_equal(te_x, np.trunc(te_x)) and np.abs(te_x).max() <= 5              # int32 truncation quirk (B-4)
    pred, err = fb.train_and_evaluate(cfg, tr_i, tr_x, te_i, te_x)
    out = capsys.readouterr().out
    assert "Test Error Rate: " in out and pred.shape == (300,) and set(np.unique(pred)) <= {0, 1} and 0 <= err <= 1
    o, w, b = fb.build_network(te_x.astype(np.float32), cfg)
    assert o.shape == (300, 2) and (o >= 0).all() and w["out"].shape == (30, 2) and set(w) == {"h1", "h2", "h3", "h4", "out"}


def test_rnn_pipeline(capsys):
    import rnn_viterbi as rv
    from oracle import lstm as olstm
    cfg = rv.RNNDecoderConfig(block_length=20, training_size=200, test_size=64, training_epochs=1, batch_size=50)
    tr_i, tr_x, te_i, te_x = rv.generate_datasets(cfg)
    assert tr_x.shape == (200, 40) and set(np.unique(te_i)) <= {-1.0, 1.0}
    model = rv.build_seq2seq_model(cfg, seed=1)
    feed = rv.prepare_batch(te_i, te_x, cfg)
    dec_in = np.transpose(feed["decoder_inputs"][:, :, 0])
    logits, hard = rv.decode(cfg, model["params"], te_x.astype(np.float32), dec_in)
    want = olstm.seq2seq_forward(te_x.astype(np.float64), dec_in, model["params"])
    assert np.abs(logits - want).max() <= 2e-5
    pred, ber = rv.train_seq2seq(cfg, tr_i, tr_x, te_i, te_x, model)
    assert pred.shape == (64, 20) and 0 <= ber <= 1 and "Bit Error Rate: " in capsys.readouterr().out


def test_production_wrapper_runs(tmp_path):
    import production_wrapper as pw
    cfg = tmp_path / "config.yaml"
    cfg.write_text(f"""
model: {{block_length: 7, snr_db: 3.0}}
training: {{epochs: 2, batch_size: 128, learning_rate: 0.001, samples_per_class_train: 10, samples_per_class_test: 10,
            checkpoint_every: 1, early_stopping_enabled: true, early_stopping_patience: 5}}
paths: {{logs_dir: {tmp_path}/logs, checkpoints_dir: {tmp_path}/ckpt, models_dir: {tmp_path}/models, results_dir: {tmp_path}/results}}
log_to_console: false
""")
    t = pw.ProductionTrainer(str(cfg))
    assert t.run() == 0
    assert (tmp_path / "results" / "training_history.json").exists() and (tmp_path / "models" / "final_model.npz").exists()
    assert len(t.metrics_tracker.history) == 2 and 0 <= t.final_metrics.ber <= 1
    assert t.checkpoint_manager.get_best_checkpoint("ber") is not None


def test_dropout_twin_runs_on_the_gpu():
    """multilayer_perceptron_with_dropout (/root/reference/src/fnn_viterbi_mixed_snr.py:596-627): keep_prob = 1 equals the
    plain graph; keep_prob < 1 applies inverted dropout after each hidden layer (reproducible from the generator)."""
    import fnn_viterbi_mixed_snr as fm
    from oracle import mlp as omlp
    cfg = fm.MixedSNRDecoderConfig()
    W, b = fm.initialize_weights(cfg)
    x = np.random.default_rng(0).standard_normal((300, 14)).astype(np.float32)
    assert np.array_equal(fm.multilayer_perceptron_with_dropout(x, W, b, 1.0), fm.multilayer_perceptron(x, W, b))
    a = fm.multilayer_perceptron_with_dropout(x, W, b, 0.5, rng=np.random.default_rng(3))
    rng = np.random.default_rng(3)
    h = x
    for wk, bk, last in (("h1", "b1", False), ("h2", "b2", False), ("h3", "b3", False), ("out", "out", True)):
        h = np.tanh(h.astype(np.float64) @ W[wk] + b[bk]).astype(np.float32)
        if not last:
            h = h * (rng.random(h.shape) < 0.5).astype(np.float32) / np.float32(0.5)
    assert np.abs(a - h).max() <= 1e-5
