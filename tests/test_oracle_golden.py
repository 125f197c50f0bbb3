"""Pin the CPU oracle against vectors computed by the LIVE reference (tools/gen_golden.py)
and against the reference tests' own known answers (/root/reference/tests/*.py)."""
import numpy as np
import pytest

from oracle import channel, encoders, lstm, metrics


def test_codebooks_match_reference(golden):
    cb = encoders.generate_class_bits(7)
    assert (cb == golden["class_bits_7"]).all()
    assert (encoders.generate_class_bits(3) == golden["class_bits_3"]).all()
    assert (encoders.encode_133171(cb) == golden["enc133171_book7"]).all()
    assert (encoders.encode_57(cb) == golden["enc57_book7"]).all()


@pytest.mark.parametrize("L", [6, 10, 33, 100, 518])
def test_random_frames_match_reference(golden, L):
    x = golden[f"enc_in_{L}"]
    assert (encoders.encode_133171(x) == golden[f"enc133171_out_{L}"]).all()
    assert (encoders.encode_57(x) == golden[f"enc57_out_{L}"]).all()
    # packed wire format round trip
    p = encoders.pack_bits(x)
    assert p.shape == (x.shape[0], (L + 7) // 8)
    cp = encoders.conv_encode_packed(p, L, encoders.CODE_133171)
    assert (encoders.unpack_bits(cp, 2 * L) == golden[f"enc133171_out_{L}"]).all()


def test_state_prefixed_encoder(golden):
    x = golden["state_in"]
    for s in range(64):
        assert (encoders.encode_133171_with_state(x, s) == golden["state_out"][s]).all()
        assert (encoders.extract_state_bits(s, 6) == golden["state_bits"][s]).all()


def test_float_inputs(golden):
    assert (encoders.encode_133171(golden["enc_float_in"]) == golden["enc_float_out"]).all()


def test_survey_appendix_a1_vectors():
    # SURVEY.md Appendix A.1 (computed from the reference while surveying)
    e = encoders.encode_133171
    assert e([1, 0, 1, 1, 0, 1, 0, 1, 1, 0]).tolist() == [1, 1, 0, 1, 0, 0, 0, 1, 1, 0, 0, 1, 0, 1, 1, 0, 0, 0, 0, 1]
    assert e([1, 0, 0, 0, 0, 0, 0, 0, 0, 0]).tolist() == [1, 1, 0, 1, 1, 1, 1, 1, 0, 0, 1, 0, 1, 1, 0, 0, 0, 0, 0, 0]
    assert encoders.encode_57([1, 1, 0, 1, 0, 0, 1]).tolist() == [1, 0, 1, 0, 1, 0, 0, 0, 0, 1, 1, 1, 1, 1]
    assert encoders.encode_133171_with_state([1, 0, 1], 1).tolist() == [0] * 10 + [1, 1, 1, 0, 1, 0, 1, 1]


def test_reference_test_suite_known_answers():
    # /root/reference/tests/test_encoders.py:44-49,75-81,128-142,169-182
    rng = np.random.default_rng(0)
    for L in (7, 10, 100):
        assert len(encoders.encode_133171(np.zeros(L, int))) == 2 * L
        assert not encoders.encode_133171(np.zeros(L, int)).any()
        assert not encoders.encode_57(np.zeros(L, int)).any()
        a, b = rng.integers(0, 2, L), rng.integers(0, 2, L)
        for enc in (encoders.encode_133171, lambda v: encoders.conv_encode(v, encoders.CODE_57_CAUSAL)):
            assert ((enc(a) ^ enc(b)) == enc(a ^ b)).all()          # GF(2) linearity
    assert encoders.extract_state_bits(5, 6).tolist() == [0, 0, 0, 1, 0, 1]
    assert len(encoders.encode_133171_with_state(np.ones(9, int), 3)) == 2 * (9 + 6)
    # test_encoders.py:184-209 expects short inputs to work (the reference raises; B-5)
    assert len(encoders.encode_133171([1])) == 2 and len(encoders.encode_133171([0] * 5)) == 10
    # test_encoders.py:227-235: state 100 must not raise (B-6)
    assert len(encoders.encode_133171_with_state([1, 0, 1], 100)) == 18
    # /root/reference/tests/test_utilities.py:72-87,193-199,236-259
    assert encoders.generate_class_bits(3)[5].tolist() == [1, 0, 1]
    assert metrics.extract_bits_single(5, 5).tolist() == [0, 0, 1, 0, 1]
    assert metrics.hamming_decimal(5, 3, 3) == 2 and metrics.hamming_decimal(0, 15, 4) == 4


def test_sigma_tables(golden):
    assert channel.snr_db_to_sigma(2.0) == pytest.approx(float(golden["sigma_2db"]), rel=1e-15)
    assert np.allclose(channel.snr_db_to_sigma(golden["snr_db_vec_mixed"]), golden["sigma_vec_mixed"], rtol=1e-15)
    assert np.allclose(channel.snr_db_to_sigma(np.arange(-2, 9)), golden["sigma_vec_bitwise"], rtol=1e-15)
    assert np.allclose(channel.snr_db_to_sigma(np.arange(0, 9)), golden["sigma_vec_rnn"], rtol=1e-15)


def test_channel_matches_reference_with_seeded_legacy_rng(golden):
    cw = golden["awgn_in"]
    np.random.seed(1234)
    y = channel.modulate_awgn(cw, 0.5617)
    assert y.dtype == np.float64 and np.array_equal(y, golden["awgn_out_seed1234"])
    np.random.seed(1234)
    y = channel.modulate_awgn(cw, 0.8902)
    assert np.array_equal(np.trunc(y).astype(np.int32), golden["awgn_out_i32_seed1234"])
    assert (channel.bpsk([0, 1, 1, 0]) == np.array([-1, 1, 1, -1], np.float32)).all()


def test_label_and_metric_helpers(golden):
    for v, bits, back in zip(golden["ebs_in"], golden["ebs_out"], golden["b2d_out"]):
        assert (metrics.extract_bits_single(int(v), 7) == bits).all()
        assert metrics.binary_to_decimal(bits) == back                 # LSB-first quirk (B-7)
        assert encoders.class_to_packed(np.array([v]), 7)[0] == back   # == time-packed word
    a, b = golden["ham_a"], golden["ham_b"]
    assert [metrics.hamming_decimal(x, y) for x, y in zip(a, b)] == golden["ham_out"].tolist()
    assert metrics.calculate_ber(a, b, 7) == pytest.approx(float(golden["ber_out"]), rel=1e-15)
    assert metrics.calculate_accuracy(a, b) == pytest.approx(float(golden["acc_out"]), rel=1e-15)
    be, fe = metrics.count_errors(a, b, 7)
    assert be == int(golden["ham_out"].sum()) and fe == int((a != b).sum())


def test_calculate_error_lines(golden):
    labels, pred = golden["calc_error_labels"], golden["calc_error_pred"]
    onehot = np.eye(16, dtype=np.float32)[labels]
    out = metrics.calculate_error(onehot, pred, encoders.generate_class_bits(4), 4, samples_per_class_test=5)
    assert "\n".join(out["lines"]) + "\n" == str(golden["calc_error_stdout"])
    # popcount form used by the CUDA reduction agrees
    be, fe = metrics.count_errors(pred, labels, 4)
    assert (be, fe) == (out["bit_errors"], out["frame_errors"])


def test_rnn_batch_layout(golden):
    enc, dec_in, dec_tgt = lstm.prepare_batch(golden["pb_info"], golden["pb_enc"], 5)
    assert np.array_equal(enc.T[:, :, None], golden["pb_encoder_inputs"].astype(np.float64))
    assert np.array_equal(dec_in.T[:, :, None], golden["pb_decoder_inputs"])
    assert np.array_equal(dec_tgt.T[:, :, None], golden["pb_decoder_targets"])


def test_lstm_matches_torch_lstm():
    torch = pytest.importorskip("torch")
    H, B, T = 20, 3, 12
    p = lstm.init_params(H, seed=3)
    p["enc_b"] = np.random.default_rng(1).standard_normal(4 * H).astype(np.float32) * 0.1
    x = np.random.default_rng(2).standard_normal((B, T))
    ours = lstm.seq2seq_forward(x, np.zeros((B, 1)), p)  # only to get encoder state indirectly
    # map TF (i,j,f,o) -> torch (i,f,g,o); +1 forget bias
    def perm(m):
        i, j, f, o = np.split(m, 4, axis=-1)
        return np.concatenate([i, f, j, o], axis=-1)
    m = torch.nn.LSTM(1, H, batch_first=True).double()
    W = perm(p["enc_W"].astype(np.float64))
    b = perm(p["enc_b"].astype(np.float64)).copy()
    b[H:2 * H] += 1.0
    with torch.no_grad():
        m.weight_ih_l0.copy_(torch.from_numpy(W[:1].T.copy()))
        m.weight_hh_l0.copy_(torch.from_numpy(W[1:].T.copy()))
        m.bias_ih_l0.copy_(torch.from_numpy(b))
        m.bias_hh_l0.zero_()
        _, (hT, cT) = m(torch.from_numpy(x[:, :, None]))
    c = np.zeros((B, H)); h = np.zeros((B, H))
    for t in range(T):
        c, h = lstm.lstm_cell_step(x[:, t:t + 1], c, h, p["enc_W"].astype(np.float64), p["enc_b"].astype(np.float64))
    assert np.allclose(h, hT[0].numpy(), atol=1e-12) and np.allclose(c, cT[0].numpy(), atol=1e-12)
    assert ours.shape == (B, 1)


def test_window_extraction(golden):
    # get_single_instance(8, coded, info, 6): window coded[6:14], label = info[3]
    assert golden["gsi_x"].tolist() == list(range(6, 14))
    assert golden["gsi_y"].tolist() == [0.0, 1.0]


# ---- decoder forward: the oracle against the reference's OWN graph functions (tools/gen_golden_forward.py) ----
import golden_models as gm  # noqa: E402
from oracle import mlp as omlp  # noqa: E402


@pytest.mark.parametrize("name,builder,x_key", [("fnn", gm.fnn, "fnn_x"), ("mixed", gm.fnn, "mixed_x"),
                                                ("bitwise", gm.bitwise, "bitwise_x"), ("seq", gm.seq_predict, "seq_x")])
def test_mlp_forward_matches_reference_graph(golden_fwd, name, builder, x_key):
    """multilayer_perceptron / build_network of the reference (run under the numpy tf stand-in, fp32) against
    oracle.mlp.forward with the same parameters: soft outputs within 1e-5, decisions equal off ties."""
    Ws, bs, acts = builder(int(golden_fwd[f"{name}_seed"]))
    x = golden_fwd[x_key].astype(np.float32)
    ref = golden_fwd[f"{name}_out"]
    got = omlp.forward(x, Ws, bs, acts, dtype=np.float64)
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max())
    ok = omlp.top2_margin(got) >= 1e-5
    assert (omlp.argmax_first(got)[ok] == golden_fwd[f"{name}_pred"][ok]).all()
    got32 = omlp.forward(x, Ws, bs, acts, dtype=np.float32)
    assert np.abs(got32 - ref).max() <= 2e-6 * max(1.0, np.abs(ref).max())


def test_bitwise_graph_shape_quirk(golden_fwd):
    # Wout is [n_hidden_1, 2]: the h2..h4 layers are created but not wired (src/fnn_viterbi_bitwise.py:371-397)
    assert golden_fwd["bitwise_wout_shape"].tolist() == [30, 2]
    assert (golden_fwd["bitwise_out"] >= 0).all()               # ReLU on the logits
    assert np.abs(golden_fwd["bitwise_x"]).max() <= 4           # int32-truncated channel output (:309-317)


def test_seq_windows_match_reference(golden_fwd):
    # get_single_instance (src/fnn_viterbi_seq_predict.py:293-336): window i = stream[2i : 2i + 200], label = info[i]
    s, info = golden_fwd["seq_stream"], golden_fwd["seq_info"]
    for i in range(golden_fwd["seq_x"].shape[0]):
        assert np.array_equal(golden_fwd["seq_x"][i], s[2 * i:2 * i + 200].astype(np.float32))
        assert golden_fwd["seq_labels_onehot"][i].argmax() == info[i]


def test_lstm_forward_matches_reference_wiring(golden_fwd):
    """build_seq2seq_model's wiring (encoder over y, decoder from the encoder's final state over the teacher-forced
    inputs, linear head) run by the reference's own code; the cell arithmetic is the stand-in's TF-1.x LSTMCell."""
    p = gm.rnn(int(golden_fwd["rnn_seed"]))
    bits, enc = golden_fwd["rnn_bits"].astype(np.float64), golden_fwd["rnn_enc"].astype(np.float64)
    _, dec_in, dec_tgt = lstm.prepare_batch(2.0 * bits - 1.0, enc, 100)
    got = lstm.seq2seq_forward(enc, dec_in, p, dtype=np.float64)
    assert np.abs(got - golden_fwd["rnn_logits"]).max() <= 1e-5
    assert float(np.mean((dec_tgt - got) ** 2)) == pytest.approx(float(golden_fwd["rnn_loss"]), rel=1e-5)
    assert not golden_fwd["rnn_prediction"].any()               # argmax over a size-1 axis (B-12)
