"""GPU parity tests proper (run with `-m gpu` on a B200): every kernel is called through the C ABI
(libnvd.so via ctypes) and compared with the CPU oracle on the same seeded inputs.
Integer / bit work is bit-exact; floating point carries its tolerance in the test."""
import numpy as np
import pytest

from oracle import channel as ochan, encoders as oenc, lstm as olstm, metrics as omet, mlp as omlp, philox as ophi

pytestmark = pytest.mark.gpu

SEED = 20170710


@pytest.fixture(scope="module")
def sim():
    from nvd import montecarlo
    with montecarlo.Simulator(device=0, seed=SEED) as s:
        yield s


def _canon_image(M, rows_by_k):
    """[rows, 16] bf16 (uint16) matrix -> UMMA canonical K-major no-swizzle image (LBO = rows*16, SBO = 128)."""
    rows = rows_by_k.shape[0]
    img = np.zeros(rows * 16, dtype=np.uint16)
    for k in range(16):
        off = (k // 8) * (rows * 8) + np.arange(rows) * 8 + (k % 8)
        img[off] = rows_by_k[:, k]
    return img


@pytest.mark.parametrize("N", [32, 64, 128])
def test_umma_layout_probe(sim, N):
    """Pins the shared-memory descriptor/layout convention the tensor-core kernel relies on."""
    import ctypes as C
    from nvd import montecarlo as mc, _lib
    rng = np.random.default_rng(N)
    A = mc.f32_to_bf16(rng.standard_normal((128, 16)).astype(np.float32))
    Bm = mc.f32_to_bf16(rng.standard_normal((N, 16)).astype(np.float32))
    want = mc.bf16_to_f32(A).astype(np.float64) @ mc.bf16_to_f32(Bm).astype(np.float64).T
    a_img, b_img = _canon_image(128, A), _canon_image(N, Bm)
    d = np.zeros((128, N), np.float32)
    _lib.check(sim.lib.nvd_debug_umma_probe(sim.ctx, N, 128 * 16, 128, N * 16, 128, a_img.ctypes.data,
                                            b_img.ctypes.data, d.ctypes.data))
    err = np.abs(d - want).max()
    if err > 1e-3:  # diagnose: try LBO/SBO swapped so a failure tells us which convention the HW uses
        d2 = np.zeros_like(d)
        sim.lib.nvd_debug_umma_probe(sim.ctx, N, 128, 128 * 16, 128, N * 16, a_img.ctypes.data, b_img.ctypes.data,
                                     d2.ctypes.data)
        pytest.fail(f"UMMA layout mismatch: max err {err}; with LBO/SBO swapped {np.abs(d2 - want).max()}")


@pytest.mark.parametrize("N", [32, 64])
def test_umma_ts_mode_probe(sim, N):
    """A operand from tensor memory (tcgen05.st by the row-owning thread, two bf16 per 32-bit column)."""
    from nvd import montecarlo as mc, _lib
    rng = np.random.default_rng(100 + N)
    A = mc.f32_to_bf16(rng.standard_normal((128, 16)).astype(np.float32))
    Bm = mc.f32_to_bf16(rng.standard_normal((N, 16)).astype(np.float32))
    want = mc.bf16_to_f32(A).astype(np.float64) @ mc.bf16_to_f32(Bm).astype(np.float64).T
    a_rows = (A[:, 0::2].astype(np.uint32) | (A[:, 1::2].astype(np.uint32) << 16)).astype(np.uint32)   # even k in the low half
    b_img = _canon_image(N, Bm)
    d = np.zeros((128, N), np.float32)
    _lib.check(sim.lib.nvd_debug_umma_probe_ts(sim.ctx, N, N * 16, 128, np.ascontiguousarray(a_rows).ctypes.data,
                                               b_img.ctypes.data, d.ctypes.data))
    assert np.abs(d - want).max() < 1e-3, np.abs(d - want).max()


# ------------------------------------------------------------------------------------------ K0 / K1
def test_source_small_and_philox(sim):
    B = 1000
    got = sim.source(7, 5, B).download(np.uint8, (B,))
    want = oenc.class_to_packed((np.arange(5, 5 + B) & 127), 7)
    assert (got == want).all()
    got = sim.source(12, 0, 4099).download(np.uint16, (4099,))
    assert (got == oenc.class_to_packed(np.arange(4099) & 4095, 12)).all()
    L, B = 518, 300
    got = sim.source(L, 77, B).download(np.uint8, (B, (L + 7) // 8))
    want = oenc.pack_bits(ophi.info_bits(SEED, np.arange(77, 77 + B, dtype=np.uint64), L))
    assert (got == want).all()


@pytest.mark.parametrize("code", [oenc.CODE_133171, oenc.CODE_57_REF, oenc.CODE_57_CAUSAL])
def test_encoder_exhaustive_L7(sim, code, golden):
    cb = oenc.generate_class_bits(7)
    got = sim.encode_bits(cb, code)
    assert (got == oenc.conv_encode(cb, code)).all()
    if code == oenc.CODE_133171:
        assert (got == golden["enc133171_book7"]).all()
    if code == oenc.CODE_57_REF:
        assert (got == golden["enc57_book7"]).all()


@pytest.mark.parametrize("L", [1, 2, 5, 6, 8, 10, 16, 33, 100, 518, 1000])
@pytest.mark.parametrize("B", [1, 17, 4099])
def test_encoder_random_frames(sim, L, B):
    rng = np.random.default_rng(L * 100003 + B)
    bits = rng.integers(0, 2, size=(B, L))
    for code in (oenc.CODE_133171, oenc.CODE_57_REF):
        if code == oenc.CODE_57_REF and L < 2:
            continue
        assert (sim.encode_bits(bits, code) == oenc.conv_encode(bits, code)).all()


def test_encoder_golden_and_state(sim, golden):
    for L in (6, 10, 33, 100, 518):
        assert (sim.encode_bits(golden[f"enc_in_{L}"], "133171") == golden[f"enc133171_out_{L}"]).all()
        assert (sim.encode_bits(golden[f"enc_in_{L}"], "57") == golden[f"enc57_out_{L}"]).all()
    x = golden["state_in"]
    pre = np.stack([np.concatenate([oenc.extract_state_bits(s), x]) for s in range(64)])
    assert (sim.encode_bits(pre, "133171") == golden["state_out"]).all()


def test_encoder_empty_and_linearity_at_scale(sim):
    from nvd import _lib
    _lib.check(sim.lib.nvd_encode(sim.ctx, 1, None, 7, 0, None))            # empty batch is a no-op
    B, L = 1 << 20, 7
    rng = np.random.default_rng(1)
    a = rng.integers(0, 128, size=B, dtype=np.uint8)
    b = rng.integers(0, 128, size=B, dtype=np.uint8)
    enc = lambda w: sim.encode("133171", sim.to_device(w), L, B).download(np.uint16, (B,))
    assert (enc(a) ^ enc(b) == enc(a ^ b)).all()                             # GF(2) linearity, 2^20 frames
    assert not enc(np.zeros(B, np.uint8)).any()


# ------------------------------------------------------------------------------------------ K2
@pytest.mark.parametrize("n,B,frame0,snr_idx", [(14, 4096, 0, 0), (14, 1001, 123456789012, 5), (200, 257, 7, 2),
                                               (15, 333, 1, 1), (1036, 64, 0, 3), (16, 500, 3, 1), (6, 77, 0, 0), (2, 33, 9, 4)])
def test_channel_matches_oracle(sim, n, B, frame0, snr_idx):
    rng = np.random.default_rng(n + B)
    coded = rng.integers(0, 2, size=(B, n))
    sigma = float(ochan.snr_db_to_sigma(1.0))
    want = ochan.channel_philox(coded, sigma, SEED, snr_idx, frame0, ochan.OUT_F32)
    got = sim.channel_bits(coded, sigma, snr_idx, frame0, "f32")
    # MUFU lg2/sin/cos approximations vs float64 libm: tolerance 1e-3 absolute, mean error far lower
    assert np.abs(got - want).max() < 1e-3
    assert np.abs(got - want).mean() < 2e-6
    got16 = sim.channel_bits(coded, sigma, snr_idx, frame0, "bf16")
    assert np.abs(got16 - want).max() < 2e-2 * (1 + np.abs(want).max())
    assert (got16 == ochan.bf16_round(got)).all()          # same device function, RN conversion
    goti = sim.channel_bits(coded, sigma, snr_idx, frame0, "i32_trunc")
    near_int = np.abs(want - np.round(want)) < 2e-3
    assert ((goti == np.trunc(want).astype(np.int32)) | near_int).all()


def test_channel_far_tail_refinement_matches_oracle(sim):
    """Radius cells k < 16 (probability 1.5e-5 per pair) take 32 more bits from a second Philox call that only the
    threads meeting such a word make: find frames with such words by the oracle's generator and hold the frame kernel
    (n = 14), the block kernel (n = 40) and its ragged tail to the oracle exactly there."""
    from oracle import philox
    snr_idx, B = 2, 1 << 19
    frames = np.arange(B, dtype=np.uint64)
    sw = (snr_idx << 8) | philox.STREAM_NOISE
    hits = np.zeros(B, bool)
    for blk in range(5):                                        # n = 40 symbols = 5 blocks of 8
        for w in philox.frame_words(SEED, sw, frames, blk):
            hits |= (w >> np.uint32(12)) < philox.TAIL_CELLS
    rows = np.flatnonzero(hits)
    assert len(rows) >= 20                                      # ~ 2^19 * 20 pairs * 1.5e-5 = 160 expected
    sigma = 0.8
    for n in (14, 40, 37):
        worst = 0.0
        for f in rows[:24]:
            f0 = int(f) - 3                                     # the hit sits inside a small batch, not at its edge
            if f0 < 0:
                continue
            coded = ((np.arange(8 * n).reshape(8, n) * 7 + int(f)) % 3 == 0).astype(np.uint8)
            want = ochan.channel_philox(coded, sigma, SEED, snr_idx, f0, ochan.OUT_F32)
            got = sim.channel_bits(coded, sigma, snr_idx, f0, "f32")
            worst = max(worst, float(np.abs(got - want).max()))
        assert worst < 1e-3, (n, worst)
    # and the tail is really there: the largest radius among the hit frames exceeds what a bare 20-bit field could give
    z = philox.noise_normals(SEED, snr_idx, rows[:200].astype(np.uint64), 40)
    assert np.sqrt(z[:, 0::2] ** 2 + z[:, 1::2] ** 2).max() > 4.7


def test_channel_moments_and_bpsk(sim):
    B, n = 1 << 16, 14
    zeros = np.zeros((B, n), np.uint8)
    y0 = sim.channel_bits(zeros, 0.5, 0, 0, "f32")
    assert abs(y0.mean() + 1.0) < 5e-3 and abs(y0.std() - 0.5) < 5e-3
    y_clean = sim.channel_bits(np.tile(np.array([0, 1] * 7, np.uint8), (16, 1)), 0.0, 0, 0, "f32")
    assert (y_clean == np.tile(np.array([-1, 1] * 7, np.float32), (16, 1))).all()       # 0 -> -1, 1 -> +1


# ------------------------------------------------------------------------------------------ K3 fp32
def _model(dims, acts, seed, bias=0.05):
    from nvd import weights
    Ws, bs = omlp.init_fnn(dims, seed=seed, bias_scale=bias)
    return weights.MlpModel(Ws, bs, acts)


@pytest.mark.parametrize("dims,acts,B", [([14, 128, 64, 32, 128], [1, 1, 1, 1], 5000),
                                         ([1036, 30, 2], [2, 2], 777),
                                         ([200, 500, 500, 500, 500, 2], [1, 1, 1, 1, 1], 300),
                                         ([14, 128], [0], 129)])
def test_mlp_fp32_matches_oracle(sim, dims, acts, B):
    m = _model(dims, acts, seed=len(dims))
    h = sim.load_mlp(m)
    rng = np.random.default_rng(B)
    y = rng.standard_normal((B, dims[0])).astype(np.float32)
    pred, soft = sim.decode(h, y, "fp32")
    want = omlp.forward(y, m.Ws, m.bs, m.acts, dtype=np.float64)
    # fp32 path: 1e-5 relative (outputs are O(1); relative to max(1, |x|))
    assert np.abs(soft - want).max() <= 1e-5 * max(1.0, np.abs(want).max())
    margin = omlp.top2_margin(want)
    ok = margin >= 1e-3
    assert (pred[ok] == omlp.argmax_first(want)[ok]).all()
    assert (pred == omlp.argmax_first(soft)).all()          # first index on ties, incl. relu (0,0) rows


def test_mlp_fp32_ml_codebook_self_test(sim):
    """'Perfect weights': one identity layer with the +-1 code book == brute-force ML decoding."""
    from nvd import weights
    Ws, bs, acts = omlp.ml_codebook_model(7)
    h = sim.load_mlp(weights.MlpModel(Ws, bs, acts))
    cb = oenc.generate_class_bits(7)
    y = (2.0 * oenc.encode_133171(cb) - 1.0).astype(np.float32)
    pred, _ = sim.decode(h, y, "fp32")
    assert (pred == np.arange(128)).all()


# ------------------------------------------------------------------------------------------ K3 tcgen05
@pytest.mark.parametrize("B", [1, 127, 128, 129, 1000, 65536 + 3])
def test_mlp_tensor_core_matches_oracle(sim, B):
    from nvd import montecarlo as mc
    m = _model([14, 128, 64, 32, 128], [1, 1, 1, 1], seed=11, bias=0.1)
    h = sim.load_mlp(m)
    assert sim.mlp_tc_supported(h)
    rng = np.random.default_rng(B)
    cls = rng.integers(0, 128, B)
    y = (2.0 * oenc.encode_133171(oenc.generate_class_bits(7))[cls] - 1.0) + 0.6 * rng.standard_normal((B, 14))
    y16 = mc.f32_to_bf16(y.astype(np.float32))
    d_y = sim.to_device(y16)
    soft_b, hard_b = sim.mlp_forward(h, d_y, "bf16", B, 128, "bf16", want_soft=True, hard_bytes=4)
    soft = soft_b.download(np.float32, (B, 128))
    pred = hard_b.download(np.int32, (B,))
    yq = mc.bf16_to_f32(y16)
    want64 = omlp.forward(yq, m.Ws, m.bs, m.acts, dtype=np.float64)
    from nvd import _lib
    model16, pre16 = omlp.forward_bf16_model(yq, m.Ws, m.bs, m.acts, tanh_model=_lib.tc_tanh_model())
    assert np.abs(soft - want64).max() <= 2e-2            # north-star bf16 tolerance vs the fp64 oracle
    assert np.abs(soft - model16).max() <= 1e-2           # vs the bf16 numerical model (two-pipe tanh, nvd_tc_tanh_model)
    ok = omlp.top2_margin(want64) >= 5e-2
    assert (pred[ok] == omlp.argmax_first(want64)[ok]).all()
    assert (pred == omlp.argmax_first(soft)).mean() > 0.999   # argmax is taken on the pre-activation


def test_mlp_tensor_core_unsupported_shape_fails_loudly(sim):
    from nvd import NvdError
    for dims in ([14, 600, 128], [1036, 100, 2], [14, 64, 256]):    # a layer wider than 511 / more than 128 classes
        m = _model(dims, [1, 1], seed=3)
        h = sim.load_mlp(m)
        assert not sim.mlp_tc_supported(h)
        with pytest.raises(NvdError):
            sim.mlp_forward(h, sim.zeros(dims[0] * 2 * 8), "bf16", 8, dims[-1], "bf16")
    # the fused Monte-Carlo step exists for the reference FNN shape only
    m = _model([14, 64, 128], [1, 1], seed=3)
    h = sim.load_mlp(m)
    assert sim.mlp_tc_supported(h)                                   # wide-layer path
    with pytest.raises(NvdError):
        sim.run(h, "133171", 7, [2.0], 1000, 0, mode="fused")


# ------------------------------------------------------------------------------------------ K5
def test_count_errors(sim):
    from nvd import _lib
    rng = np.random.default_rng(5)
    B, L = 100003, 7
    label_cls = rng.integers(0, 128, B)
    pred_cls = np.where(rng.random(B) < 0.1, rng.integers(0, 128, B), label_cls)
    be, fe = omet.count_errors(pred_cls, label_cls, L)
    counters = sim.zeros(32)
    d_pred = sim.to_device(pred_cls.astype(np.uint8))
    d_lab = sim.to_device(oenc.class_to_packed(label_cls, L).astype(np.uint8))
    sim.count_errors(d_pred, d_lab, 1, L, B, _lib.COUNT_PRED_IS_MSB_CLASS, counters)
    assert counters.download(np.uint64, (4,)).tolist() == [be, fe, B, B * L]
    sim.count_errors(d_pred, d_lab, 1, L, B, _lib.COUNT_PRED_IS_MSB_CLASS, counters)      # accumulates
    assert counters.download(np.uint64, (4,)).tolist() == [2 * be, 2 * fe, 2 * B, 2 * B * L]
    # packed rows (LSTM-style, L = 100 -> 13 bytes), no class flag
    Lr, Br = 100, 5001
    a = rng.integers(0, 2, (Br, Lr)); b = a ^ (rng.random((Br, Lr)) < 0.01)
    c2 = sim.zeros(32)
    sim.count_errors(sim.to_device(oenc.pack_bits(a)), sim.to_device(oenc.pack_bits(b)), 13, Lr, Br, 0, c2)
    assert c2.download(np.uint64, (4,)).tolist() == [int((a != b).sum()), int((a != b).any(1).sum()), Br, Br * Lr]
    # 16-bit classes
    L16 = 12
    lab = rng.integers(0, 4096, 999); prd = lab ^ (rng.integers(0, 4096, 999) * (rng.random(999) < 0.2))
    c3 = sim.zeros(32)
    sim.count_errors(sim.to_device(prd.astype(np.uint16)), sim.to_device(oenc.class_to_packed(lab, L16).astype(np.uint16)),
                     2, L16, 999, _lib.COUNT_PRED_IS_MSB_CLASS, c3)
    be, fe = omet.count_errors(prd, lab, L16)
    assert c3.download(np.uint64, (4,)).tolist() == [be, fe, 999, 999 * L16]


# ------------------------------------------------------------------------------------------ K4
@pytest.mark.parametrize("teacher", [True, False])
def test_lstm_matches_oracle(sim, teacher):
    from nvd import weights
    H, L, B = 20, 100, 300
    p = olstm.init_params(H, seed=4)
    rng = np.random.default_rng(9)
    p["enc_b"] = (0.1 * rng.standard_normal(4 * H)).astype(np.float32)
    p["dec_b"] = (0.1 * rng.standard_normal(4 * H)).astype(np.float32)
    p["lin_b"] = np.array([0.05], np.float32)
    info = rng.integers(0, 2, (B, L))
    y = (2.0 * oenc.encode_133171(info) - 1.0 + 0.6 * rng.standard_normal((B, 2 * L))).astype(np.float32)
    _, dec_in, _ = olstm.prepare_batch(2.0 * info - 1.0, y, L)
    want = olstm.seq2seq_forward(y, dec_in, p, dtype=np.float64, autoregressive=not teacher)
    h = sim.load_lstm(weights.LstmModel(**p))
    d_dec = sim.to_device(dec_in.astype(np.float32)) if teacher else None
    logits_b, hard_b = sim.lstm_forward(h, sim.to_device(y), "f32", d_dec, 2 * L, 2 * L, B, fast_math=False)
    got = logits_b.download(np.float32, (B, 2 * L))
    hard = np.unpackbits(hard_b.download(np.uint8, (B, 25)), axis=1, bitorder="little")[:, : 2 * L]
    if teacher:
        assert np.abs(got - want).max() <= 2e-5           # fp32 recurrence over 400 steps vs fp64
        ok = np.abs(want) >= 1e-3
        assert (hard[ok] == (want > 0)[ok]).all()
    else:
        same = (np.sign(got) == np.sign(want)).all(axis=1)  # a flipped near-zero logit changes the fed-back path
        assert same.mean() > 0.98
        assert np.abs(got[same] - want[same]).max() <= 2e-4
    assert (hard == (got > 0)).all()
    fast_b, _ = sim.lstm_forward(h, sim.to_device(y), "f32", d_dec, 2 * L, 2 * L, B, fast_math=True)
    if teacher:
        assert np.abs(fast_b.download(np.float32, (B, 2 * L)) - want).max() <= 2e-2   # MUFU tanh path


@pytest.mark.parametrize("H,B", [(20, 300), (16, 129), (20, 148 * 4 * 128 + 77)])
def test_lstm_tensor_core_recurrence(sim, H, B):
    """fast_math = 2: the tcgen05 recurrence (bf16 weights and h, fp32 accumulation and cell state) against the
    fp64 oracle, teacher-forced over 200 + 200 steps: bf16 tolerance, decisions equal away from zero; the packed
    hard decisions are the signs of its own logits; autoregressive mode agrees with the fp32 kernel's decisions."""
    from nvd import weights
    L = 100
    p = olstm.init_params(H, seed=4)
    rng = np.random.default_rng(9)
    p["enc_b"] = (0.1 * rng.standard_normal(4 * H)).astype(np.float32)
    p["dec_b"] = (0.1 * rng.standard_normal(4 * H)).astype(np.float32)
    p["lin_b"] = np.array([0.05], np.float32)
    info = rng.integers(0, 2, (B, L))
    y = (2.0 * oenc.encode_133171(info) - 1.0 + 0.6 * rng.standard_normal((B, 2 * L))).astype(np.float32)
    _, dec_in, _ = olstm.prepare_batch(2.0 * info - 1.0, y, L)
    h = sim.load_lstm(weights.LstmModel(**p))
    d_y, d_dec = sim.to_device(y), sim.to_device(dec_in.astype(np.float32))
    logits_b, hard_b = sim.lstm_forward(h, d_y, "f32", d_dec, 2 * L, 2 * L, B, fast_math=2)
    got = logits_b.download(np.float32, (B, 2 * L))
    hard = np.unpackbits(hard_b.download(np.uint8, (B, 25)), axis=1, bitorder="little")[:, : 2 * L]
    assert (hard == (got > 0)).all()
    n_chk = min(B, 2000)
    want = olstm.seq2seq_forward(y[:n_chk], dec_in[:n_chk], p, dtype=np.float64)
    err = np.abs(got[:n_chk] - want)
    assert err.max() <= 3e-2, err.max()                   # bf16 operands through 400 recurrent steps
    ok = np.abs(want) >= 5e-2
    assert (hard[:n_chk][ok] == (want > 0)[ok]).all()
    # the tail of the batch (last, partial tile) against the fp32 kernel
    ref_b, _ = sim.lstm_forward(h, d_y, "f32", d_dec, 2 * L, 2 * L, B, fast_math=False, want_hard=False)
    ref = ref_b.download(np.float32, (B, 2 * L))
    assert np.abs(got[-200:] - ref[-200:]).max() <= 3e-2
    # autoregressive: decisions fed back
    _, hard_a = sim.lstm_forward(h, d_y, "f32", None, 2 * L, 2 * L, B, fast_math=2, want_logits=False)
    _, hard_r = sim.lstm_forward(h, d_y, "f32", None, 2 * L, 2 * L, B, fast_math=False, want_logits=False)
    ha = np.unpackbits(hard_a.download(np.uint8, (B, 25)), axis=1, bitorder="little")[:, : 2 * L]
    hr = np.unpackbits(hard_r.download(np.uint8, (B, 25)), axis=1, bitorder="little")[:, : 2 * L]
    assert (ha == hr).all(axis=1).mean() > 0.9            # a flipped near-zero logit changes the fed-back path


# ------------------------------------------------------------------------------------------ K3-wide
@pytest.mark.parametrize("dims,acts,B", [([200, 500, 500, 500, 500, 2], [1, 1, 1, 1, 1], 1000),   # seq_predict (:343-414)
                                         ([200, 500, 500, 500, 500, 2], [1, 1, 1, 1, 1], 148 * 128 + 5),
                                         ([30, 64, 10], [2, 0], 129),                             # relu hidden, identity out
                                         ([17, 129, 70, 128], [1, 2, 1], 300),                    # chunk tails, 128 classes
                                         ([14, 20], [1], 77)])
def test_mlp_wide_tensor_core_matches_oracle(sim, dims, acts, B):
    """Layer-per-launch tcgen05 path (A operand in tensor memory, weights streamed through a TMA ring) against the
    fp64 oracle and the bf16 numerical model, inputs in fp32 (converted on the device)."""
    m = _model(dims, acts, seed=len(dims) + dims[1], bias=0.1)
    h = sim.load_mlp(m)
    assert sim.mlp_tc_supported(h)
    rng = np.random.default_rng(B)
    y = rng.standard_normal((B, dims[0])).astype(np.float32)
    soft_b, hard_b = sim.mlp_forward(h, sim.to_device(y), "f32", B, dims[-1], "bf16", want_soft=True, hard_bytes=4)
    soft = soft_b.download(np.float32, (B, dims[-1]))
    pred = hard_b.download(np.int32, (B,))
    from nvd import montecarlo as mc
    yq = mc.bf16_to_f32(mc.f32_to_bf16(y))
    want64 = omlp.forward(yq, m.Ws, m.bs, m.acts, dtype=np.float64)
    model16, _ = omlp.forward_bf16_model(yq, m.Ws, m.bs, m.acts)
    scale = max(1.0, float(np.abs(want64).max()))
    assert np.abs(soft - want64).max() <= 2e-2 * scale                # north-star bf16 tolerance
    assert np.abs(soft - model16).max() <= 1e-2 * scale               # bf16 numerical model (bf16 bias as well)
    ok = omlp.top2_margin(want64) >= 5e-2 * scale
    assert (pred[ok] == omlp.argmax_first(want64)[ok]).all()


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_window_decode_matches_dense_rows(sim, precision):
    """nvd_window_decode over a stride-2 stream == the same windows decoded as dense rows (get_single_instance,
    fnn_viterbi_seq_predict.py:293-336), both precisions, fp32 and int32-truncated streams."""
    dims, acts = [200, 500, 500, 500, 500, 2], [1, 1, 1, 1, 1]
    m = _model(dims, acts, seed=3, bias=0.05)
    h = sim.load_mlp(m)
    rng = np.random.default_rng(0)
    n_win = 3000
    n_sym = 2 * (n_win - 1) + 200
    for dtype, arr in (("f32", rng.standard_normal(n_sym).astype(np.float32)),
                       ("i32_trunc", rng.integers(-3, 4, n_sym).astype(np.int32))):
        d_stream = sim.to_device(arr)
        soft_b, hard_b = sim.window_decode(h, d_stream, dtype, n_sym, 2, n_win, 2, precision)
        soft = soft_b.download(np.float32, (n_win, 2))
        hard = hard_b.download(np.int32, (n_win,))
        rows = np.lib.stride_tricks.sliding_window_view(arr.astype(np.float32), 200)[::2][:n_win]
        s2_b, h2_b = sim.mlp_forward(h, sim.to_device(np.ascontiguousarray(rows)), "f32", n_win, 2, precision, want_soft=True, hard_bytes=4)
        assert np.array_equal(soft, s2_b.download(np.float32, (n_win, 2)))
        assert np.array_equal(hard, h2_b.download(np.int32, (n_win,)))
        want = omlp.forward(rows, m.Ws, m.bs, m.acts, dtype=np.float64)
        tol = 1e-5 if precision == "fp32" else 2e-2
        assert np.abs(soft - want).max() <= tol
    from nvd import NvdError
    with pytest.raises(NvdError):
        sim.window_decode(h, d_stream, "f32", n_sym, 2, n_win + 1, 2, precision)   # runs past the end of the stream


@pytest.mark.parametrize("H,B,T_enc,T_dec", [(20, 1, 200, 200), (20, 131, 70, 45), (16, 700, 33, 64)])
def test_lstm_tensor_core_ragged_shapes(sim, H, B, T_enc, T_dec):
    """Edge cases of the tcgen05 recurrence: a single frame, T_enc != T_dec, lengths that are not multiples of the
    32-step staging window; against the fp32 kernel (teacher-forced logits) and its own packed decisions."""
    from nvd import weights
    p = olstm.init_params(H, seed=7)
    rng = np.random.default_rng(B + T_enc)
    p["enc_b"] = (0.1 * rng.standard_normal(4 * H)).astype(np.float32)
    p["dec_b"] = (0.1 * rng.standard_normal(4 * H)).astype(np.float32)
    y = rng.standard_normal((B, T_enc)).astype(np.float32)
    dec_in = np.sign(rng.standard_normal((B, T_dec))).astype(np.float32)
    h = sim.load_lstm(weights.LstmModel(**p))
    d_y, d_dec = sim.to_device(y), sim.to_device(dec_in)
    nb = (T_dec + 7) // 8
    tc_l, tc_h = sim.lstm_forward(h, d_y, "f32", d_dec, T_enc, T_dec, B, fast_math=2)
    ref_l, _ = sim.lstm_forward(h, d_y, "f32", d_dec, T_enc, T_dec, B, fast_math=False)
    got, ref = tc_l.download(np.float32, (B, T_dec)), ref_l.download(np.float32, (B, T_dec))
    assert np.abs(got - ref).max() <= 3e-2
    want = olstm.seq2seq_forward(y, dec_in, p, dtype=np.float64)
    assert np.abs(got - want).max() <= 3e-2
    hard = np.unpackbits(tc_h.download(np.uint8, (B, nb)), axis=1, bitorder="little")[:, :T_dec]
    assert (hard == (got > 0)).all()


def test_mlp_wide_single_row_and_empty(sim):
    m = _model([200, 500, 2], [1, 1], seed=9, bias=0.1)
    h = sim.load_mlp(m)
    y = np.random.default_rng(1).standard_normal((1, 200)).astype(np.float32)
    soft_b, hard_b = sim.mlp_forward(h, sim.to_device(y), "f32", 1, 2, "bf16", want_soft=True, hard_bytes=4)
    want = omlp.forward(y, m.Ws, m.bs, m.acts, dtype=np.float64)
    assert np.abs(soft_b.download(np.float32, (1, 2)) - want).max() <= 2e-2
    sim.mlp_forward(h, sim.to_device(y), "f32", 0, 2, "bf16", want_soft=False, hard_bytes=4)      # B = 0 is a no-op


# ------------------------------------------------------------------ reference-run golden outputs (parity pins)
import golden_models as gm  # noqa: E402

_GOLDEN_NETS = [("fnn", gm.fnn, "fnn_x"), ("mixed", gm.fnn, "mixed_x"), ("bitwise", gm.bitwise, "bitwise_x"),
                ("seq", gm.seq_predict, "seq_x")]


@pytest.mark.parametrize("name,builder,x_key", _GOLDEN_NETS)
def test_mlp_fp32_matches_reference_graph(sim, golden_fwd, name, builder, x_key):
    """The fp32 CUDA path against soft outputs the REFERENCE'S OWN multilayer_perceptron / build_network produced
    (tools/gen_golden_forward.py) on reference-made noisy inputs with identical weights: 1e-5 relative,
    hard decisions identical except where the top-2 margin is below 1e-3 (north-star criteria)."""
    from nvd import weights
    Ws, bs, acts = builder(int(golden_fwd[f"{name}_seed"]))
    h = sim.load_mlp(weights.MlpModel(Ws, bs, acts))
    x = golden_fwd[x_key].astype(np.float32)
    ref, ref_pred = golden_fwd[f"{name}_out"], golden_fwd[f"{name}_pred"]
    pred, soft = sim.decode(h, x, "fp32")
    assert np.abs(soft - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max())
    ok = omlp.top2_margin(ref) >= 1e-3
    assert ok.sum() > 0.5 * len(ok) and (pred[ok] == ref_pred[ok]).all()


@pytest.mark.parametrize("name,builder,x_key", _GOLDEN_NETS)
def test_mlp_tensor_core_matches_reference_graph(sim, golden_fwd, name, builder, x_key):
    """The bf16 tcgen05 paths (fused FNN kernel, wide-layer kernel, K-streamed bitwise kernel) against the same
    reference-run outputs: 2e-2 absolute; decisions identical where the margin exceeds twice that."""
    from nvd import weights
    Ws, bs, acts = builder(int(golden_fwd[f"{name}_seed"]))
    h = sim.load_mlp(weights.MlpModel(Ws, bs, acts))
    assert sim.mlp_tc_supported(h)
    x = golden_fwd[x_key].astype(np.float32)
    ref, ref_pred = golden_fwd[f"{name}_out"], golden_fwd[f"{name}_pred"]
    pred, soft = sim.decode(h, x, "bf16")
    scale = max(1.0, np.abs(ref).max())       # ReLU logits of the bitwise net are not bounded by 1
    assert np.abs(soft - ref).max() <= 2e-2 * scale
    ok = omlp.top2_margin(ref) >= 5e-2 * scale
    assert ok.sum() > 50 and (pred[ok] == ref_pred[ok]).all()


@pytest.mark.parametrize("fast_math,tol", [(0, 2e-5), (2, 3e-2)])
def test_lstm_matches_reference_wiring(sim, golden_fwd, fast_math, tol):
    """K4 against the logits of the reference's build_seq2seq_model wiring (teacher-forced, 200 + 200 steps)."""
    from nvd import weights
    p = gm.rnn(int(golden_fwd["rnn_seed"]))
    bits, enc = golden_fwd["rnn_bits"].astype(np.float32), golden_fwd["rnn_enc"].astype(np.float32)
    B, L = bits.shape
    _, dec_in, _ = olstm.prepare_batch(2.0 * bits - 1.0, enc, L)
    h = sim.load_lstm(weights.LstmModel(**p))
    logits_b, hard_b = sim.lstm_forward(h, sim.to_device(enc), "f32", sim.to_device(dec_in.astype(np.float32)), 2 * L, 2 * L, B,
                                        fast_math=fast_math)
    got = logits_b.download(np.float32, (B, 2 * L))
    ref = golden_fwd["rnn_logits"]
    assert np.abs(got - ref).max() <= tol
    hard = np.unpackbits(hard_b.download(np.uint8, (B, 25)), axis=1, bitorder="little")[:, : 2 * L]
    ok = np.abs(ref) >= (1e-3 if fast_math == 0 else 5e-2)
    assert (hard[ok] == (ref > 0)[ok]).all()
