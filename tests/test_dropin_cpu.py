"""Host-side parts of the drop-in modules (no GPU): configs, label helpers, both module spellings.
Mirrors /root/reference/tests/test_utilities.py; the 4 assertions there that contradict the reference
code (MSB- vs LSB-first binary_to_decimal) are pinned to the CODE's behaviour."""
import importlib

import numpy as np
import pytest


@pytest.mark.parametrize("name", ["fnn_viterbi", "fnn_viterbi_improved", "fnn_viterbi_mixed_snr", "fnn_viterbi_mixed_snr_improved",
                                  "fnn_viterbi_bitwise", "fnn_viterbi_bitwise_improved", "fnn_viterbi_seq_predict", "rnn_viterbi"])
def test_modules_import_under_both_spellings(name):
    m = importlib.import_module(name)
    assert hasattr(m, "modulate_awgn")


def test_configs_match_reference(golden):
    import fnn_viterbi as fv, fnn_viterbi_mixed_snr as fm, fnn_viterbi_bitwise as fb, rnn_viterbi as rv, fnn_viterbi_seq_predict as fs
    c = fv.DecoderConfig()
    assert (c.n_input, c.n_classes, c.num_test_samples, c.num_training_samples) == (14, 128, 1280000, 64000)
    assert c.sigma == pytest.approx(float(golden["sigma_2db"]), rel=1e-15)
    m = fm.MixedSNRDecoderConfig()
    assert np.allclose(m.sigma_vec, golden["sigma_vec_mixed"], rtol=1e-15) and m.samples_per_snr == int(golden["samples_per_snr_mixed"])
    assert np.allclose(fb.BitwiseDecoderConfig().sigma_vec, golden["sigma_vec_bitwise"], rtol=1e-15)
    assert fb.BitwiseDecoderConfig().n_input == 1036
    r = rv.RNNDecoderConfig()
    assert np.allclose(r.sigma_vec, golden["sigma_vec_rnn"], rtol=1e-15) and (r.PAD, r.EOS, r.vocab_size) == (0, 1, 1)
    assert fs.SequentialDecoderConfig().total_test_instances == int(golden["seq_total_test_instances"])


def test_class_bits_and_one_hot(golden):
    import fnn_viterbi as fv, fnn_viterbi_mixed_snr as fm
    assert (fv.generate_class_bits(7) == golden["class_bits_7"]).all()
    for L in (3, 5, 7):                                  # test_utilities.py:30-96
        cb = fv.generate_class_bits(L)
        assert cb.shape == (2 ** L, L) and cb.dtype == np.uint8 and len({tuple(r) for r in cb}) == 2 ** L
    assert fv.generate_class_bits(3)[5].tolist() == [1, 0, 1]
    oh = fv.create_one_hot_labels(np.array([0, 2, 1]), 3)
    assert (oh == golden["one_hot_3"]).all() and oh.dtype == np.float32
    assert (oh.sum(1) == 1).all() and (np.argmax(oh, 1) == [0, 2, 1]).all()
    assert fm.create_one_hot_labels(np.array([1]), 4).dtype == np.float16


def test_label_helpers_keep_reference_quirks(golden):
    import fnn_viterbi_mixed_snr as fm, fnn_viterbi_bitwise as fb
    for v, bits, back in zip(golden["ebs_in"], golden["ebs_out"], golden["b2d_out"]):
        assert (fm.extract_bits_single(int(v), 7) == bits).all()
        assert fm.binary_to_decimal(bits) == back            # LSB-first, as the reference code
        assert fm.binary_to_decimal_msb(bits) == v           # what its docstring / tests describe
    assert fm.extract_bits_single(5, 5).tolist() == [0, 0, 1, 0, 1]
    assert fm.calculate_hamming_distance_decimal(5, 3, 3) == 2 and fm.calculate_hamming_distance_decimal(0, 15, 4) == 4
    for a, b, c in [(3, 9, 12), (0, 7, 5)]:                # triangle inequality, test_utilities.py:261-271
        h = fm.calculate_hamming_distance_decimal
        assert h(a, c, 4) <= h(a, b, 4) + h(b, c, 4)
    assert (fb.extract_state_bits(5, 6) == [0, 0, 0, 1, 0, 1]).all()
    for s in range(64):
        assert (fb.extract_state_bits(s, 6) == golden["state_bits"][s]).all()
    assert len(fb.extract_state_bits(100, 6)) == 6          # test_encoders.py:227-235 expects no exception
    assert fm.calculate_accuracy(golden["ham_a"], golden["ham_b"]) == pytest.approx(float(golden["acc_out"]))


def test_window_and_batch_layout(golden):
    import fnn_viterbi_seq_predict as fs, rnn_viterbi as rv
    win = np.arange(40, dtype=np.float64).reshape(1, 40)
    info = (np.arange(20) % 2).reshape(1, 20)
    x, y = fs.get_single_instance(8, win, info, 6)
    assert (x == golden["gsi_x"]).all() and (y == golden["gsi_y"]).all()
    assert (fs.windows(win[0], 8, 3, 3)[0] == x).all()
    cfg = rv.RNNDecoderConfig(block_length=5)
    pb = rv.prepare_batch(golden["pb_info"], golden["pb_enc"], cfg)
    for k in ("encoder_inputs", "decoder_inputs", "decoder_targets"):
        assert np.array_equal(pb[k], golden[f"pb_{k}"])


def test_utils_package(tmp_path):
    from utils.config import ConfigManager
    from utils.metrics import EarlyStopping
    cfgfile = tmp_path / "c.yaml"
    cfgfile.write_text("model:\n  block_length: 7\ntraining:\n  epochs: 3\n  early_stopping_patience: 2\npaths:\n  logs_dir: %s\n" % tmp_path)
    c = ConfigManager(str(cfgfile)).get()
    assert c.model.block_length == 7 and c.training.epochs == 3 and c.training.batch_size == 256 and c.logging_level == "INFO"
    es = EarlyStopping(patience=2, min_delta=0.0)
    assert [es(v) for v in (1.0, 0.9, 0.95, 0.97)] == [False, False, False, True]
