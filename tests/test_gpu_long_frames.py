"""GPU parity tests of the long-frame paths (BASELINE.json configs[2] and configs[3]) through the C ABI:
the K-streamed tensor-core kernel of the per-bit predictor (1036-30-2), its staged and fused Monte-Carlo
steps, the recurrent decoder's Monte-Carlo step, and the K0 / K5 helpers they add."""
import numpy as np
import pytest

from oracle import channel as ochan, encoders as oenc, lstm as olstm, mlp as omlp, pipeline as opipe

pytestmark = pytest.mark.gpu
SEED = 20170710


@pytest.fixture(scope="module")
def sim():
    from nvd import montecarlo
    with montecarlo.Simulator(device=0, seed=SEED) as s:
        yield s


def _model(dims, acts, seed, bias=0.05):
    from nvd import weights
    Ws, bs = omlp.init_fnn(dims, seed=seed, bias_scale=bias)
    return weights.MlpModel(Ws, bs, acts)


def test_state_prefix_and_bit_counters(sim):
    import ctypes as C
    from nvd import _lib
    rng = np.random.default_rng(1)
    for L, state, B in ((512, 0, 300), (512, 37, 77), (20, 5, 1), (100, 63, 1000)):
        info = rng.integers(0, 2, (B, L)).astype(np.uint8)
        packed = np.packbits(info, axis=1, bitorder="little")
        d_in = sim.to_device(packed)
        S_out = (L + 6 + 7) // 8
        d_out = sim.alloc(B * S_out)
        _lib.check(sim.lib.nvd_prepend_state(sim.ctx, d_in.ptr, L, state, 6, B, d_out.ptr))
        got = np.unpackbits(d_out.download(np.uint8, (B, S_out)), axis=1, bitorder="little")
        want = np.concatenate([np.tile(oenc.extract_state_bits(state, 6), (B, 1)), info], axis=1)
        assert (got[:, :L + 6] == want).all() and not got[:, L + 6:].any()
        # ... and the encoder on top of it equals encode_133171_with_state
        coded = sim.encode(1, d_out, L + 6, B).download(np.uint8, (B, (2 * (L + 6) + 7) // 8))
        cb = np.unpackbits(coded, axis=1, bitorder="little")[:, :2 * (L + 6)]
        assert (cb[0] == oenc.encode_133171_with_state(info[0], state)).all()
        pred = rng.integers(0, 2, B).astype(np.uint8)
        for pos in (0, L // 2, L - 1):
            c = sim.count_bit_errors(pred, info, pos)
            e = int((pred != info[:, pos]).sum())
            assert c.tolist() == [e, e, B, B]
        rows = rng.integers(0, 2, (B, 2 * L)).astype(np.uint8)
        c = sim.count_row_errors(rows, info)
        x = rows[:, :L] ^ info
        assert c.tolist() == [int(x.sum()), int(x.any(axis=1).sum()), B, B * L]


@pytest.mark.parametrize("dims,acts,B", [([1036, 30, 2], [2, 2], 1), ([1036, 30, 2], [2, 2], 5000), ([1040, 30, 2], [2, 2], 129),
                                         ([1038, 30, 2], [2, 2], 257), ([100, 64, 8], [1, 0], 127), ([48, 16, 2], [1, 1], 1000),
                                         ([1200, 20, 3], [2, 1], 700), ([104, 64, 8], [1, 2], 3000), ([1040, 30, 2], [2, 2], 1),
                                         ([64, 8, 2], [0, 0], 148 * 4 * 128 * 2 + 5)])
def test_kstream_tensor_core_matches_oracle(sim, dims, acts, B):
    """bf16 rows x bf16 W1 on tcgen05 with fp32 accumulation, the tail in fp32: against the fp64 oracle on the same
    bf16-rounded inputs within 2e-2 of the output scale, and tighter against the oracle with W1 rounded to bf16.
    Row strides of 2072 / 2076 bytes take the 8- / 4-byte cp.async paths, 16-byte aligned strides (1040, 104, 64, 48
    symbols) the TMA path — incl. N1 = 64 (two 64-column accumulators per group), a single row, a single-stage K and
    more tiles than two rounds of the persistent grid."""
    from nvd import montecarlo as mc
    m = _model(dims, acts, seed=sum(dims))
    h = sim.load_mlp(m)
    assert sim.mlp_tc_supported(h)
    rng = np.random.default_rng(B)
    y = np.trunc(2.0 * rng.standard_normal((B, dims[0]))).astype(np.float32) if acts[0] == 2 else rng.standard_normal((B, dims[0])).astype(np.float32)
    y16 = mc.f32_to_bf16(y)
    yq = mc.bf16_to_f32(y16)
    soft_b, hard_b = sim.mlp_forward(h, sim.to_device(y16), "bf16", B, dims[-1], "bf16", want_soft=True, hard_bytes=4)
    soft, pred = soft_b.download(np.float32, (B, dims[-1])), hard_b.download(np.int32, (B,))
    want = omlp.forward(yq, m.Ws, m.bs, m.acts, dtype=np.float64)
    Wq = [ochan.bf16_round(m.Ws[0])] + list(m.Ws[1:])
    model = omlp.forward(yq, Wq, m.bs, m.acts, dtype=np.float64)
    scale = max(1.0, np.abs(want).max())
    assert np.abs(soft - want).max() <= 2e-2 * scale
    assert np.abs(soft - model).max() <= 2e-4 * scale       # same rounding points: only the accumulation order differs
    ok = omlp.top2_margin(want) >= 5e-2 * scale
    assert (pred[ok] == omlp.argmax_first(want)[ok]).all()
    assert (pred == omlp.argmax_first(soft)).all()           # first index on ties, incl. ReLU (0, 0) rows
    # fp32 rows through the same entry (converted on the device)
    soft2, _ = sim.mlp_forward(h, sim.to_device(yq), "f32", B, dims[-1], "bf16", want_soft=True, hard_bytes=4)
    assert np.array_equal(soft2.download(np.float32, (B, dims[-1])), soft)


def test_kstream_empty_and_unsupported(sim):
    from nvd import NvdError
    m = _model([1036, 30, 2], [2, 2], seed=1)
    h = sim.load_mlp(m)
    sim.mlp_forward(h, sim.zeros(64), "bf16", 0, 2, "bf16")                 # empty batch is a no-op
    h2 = sim.load_mlp(_model([1037, 30, 2], [2, 2], seed=2))                 # odd width: rows are not 4-byte aligned
    assert not sim.mlp_tc_supported(h2)
    with pytest.raises(NvdError):
        sim.mlp_forward(h2, sim.zeros(1037 * 2 * 4), "bf16", 4, 2, "bf16")
    h3 = sim.load_mlp(_model([1036, 100, 2], [2, 2], seed=3))                # first layer wider than 64, K > 511
    assert not sim.mlp_tc_supported(h3)


@pytest.mark.parametrize("frames,frame0,trunc", [(1, 0, True), (300, 7, True), (70000, 1 << 33, True), (4097, 3, False)])
def test_bitwise_fused_equals_staged(sim, frames, frame0, trunc):
    m = _model([1036, 30, 2], [2, 2], seed=9)
    h = sim.load_mlp(m)
    snr = [0.0, 5.0]
    kw = dict(state=21, bit_position=3, trunc=trunc)
    fused = sim.run_bitwise(h, 512, snr, frames, frame0, mode="fused", **kw)
    staged = sim.run_bitwise(h, 512, snr, frames, frame0, mode="staged", **kw)
    assert (fused == staged).all()                     # bit-identical frames and arithmetic
    assert (fused[:, 2] == frames).all() and (fused[:, 3] == frames).all() and (fused[:, 0] == fused[:, 1]).all()
    # any split of the frame range gives the same counters
    from nvd.montecarlo import shard_range
    acc = np.zeros_like(fused)
    for rank in range(3):
        lo, n = shard_range(frame0, frames, rank, 3)
        acc += sim.run_bitwise(h, 512, snr, n, lo, mode="fused", **kw)
    assert (acc == fused).all()


@pytest.mark.parametrize("trunc", [True, False])
def test_bitwise_counters_match_oracle(sim, trunc):
    """Same Philox frames through the CPU oracle (fnn_viterbi_bitwise.py:290-323, :388-397, :504-509): the fp32 path
    may differ only on near-tie frames, the tensor-core path only where the margin is inside the bf16 tolerance."""
    from nvd import weights
    import os
    path = os.path.join(weights.WEIGHTS_DIR, "bitwise_133171_L512_bit0.npz")
    m = weights.load_mlp(path) if os.path.exists(path) else _model([1036, 30, 2], [2, 2], seed=9)
    h = sim.load_mlp(m)
    frames, db = 4096, 3.0
    ref, det = opipe.mc_step_bitwise(m.Ws, m.bs, m.acts, 512, db, 100, frames, seed=SEED, snr_idx=2, trunc=trunc,
                                     precision="fp64", return_detail=True)
    margin = omlp.top2_margin(det["out"])
    scale = max(1.0, np.abs(det["out"]).max())
    fp32 = sim.run_bitwise(h, 512, [db], frames, 100, mode="staged", precision="fp32", trunc=trunc, snr_idx=[2])[0]
    assert abs(int(fp32[0]) - int(ref[0])) <= int((margin < 2e-3 * scale).sum())     # channel noise itself is ~1e-3 accurate
    tc = sim.run_bitwise(h, 512, [db], frames, 100, mode="fused", precision="bf16", trunc=trunc, snr_idx=[2])[0]
    assert abs(int(tc[0]) - int(ref[0])) <= int((margin < 5e-2 * scale).sum())
    assert tc[2] == frames and tc[3] == frames
    if trunc:                                            # integer inputs are exact in bf16: only W1's rounding is left
        q = opipe.mc_step_bitwise(m.Ws, m.bs, m.acts, 512, db, 100, frames, seed=SEED, snr_idx=2, trunc=True, precision="bf16")
        assert abs(int(tc[0]) - int(q[0])) <= 4


@pytest.mark.parametrize("teacher", [True, False])
def test_lstm_monte_carlo_step_matches_oracle(sim, teacher):
    """nvd_mc_run_lstm (K0 -> K1 -> K2 -> K4 -> K5 on packed rows) against the oracle's statement of
    rnn_viterbi.py:186-232, :328-377, :278-305, :471-491 on the same Philox frames."""
    from nvd import weights
    import os
    path = os.path.join(weights.WEIGHTS_DIR, "rnn_133171_L100.npz")
    if os.path.exists(path):
        p = {k: v for k, v in np.load(path).items()}
    else:
        p = olstm.init_params(20, seed=4)
    h = sim.load_lstm(weights.LstmModel(**p))
    frames, db, L = 600, 2.0, 100
    ref, det = opipe.mc_step_lstm(p, L, db, 50, frames, seed=SEED, snr_idx=2, teacher_forced=teacher, return_detail=True)
    near = int((np.abs(det["logits"][:, :L]) < 1e-3).sum())
    got = sim.run_lstm(h, L, [db], frames, 50, fast_math=0, teacher_forced=teacher, snr_idx=[2])[0]
    assert got[2] == frames and got[3] == frames * L
    if teacher:
        assert abs(int(got[0]) - int(ref[0])) <= near + 2
    else:
        assert abs(int(got[0]) - int(ref[0])) <= 0.02 * frames * L      # a flipped near-zero logit changes the fed-back path
    tc = sim.run_lstm(h, L, [db], frames, 50, fast_math=2, teacher_forced=teacher, snr_idx=[2])[0]
    near_tc = int((np.abs(det["logits"][:, :L]) < 5e-2).sum())
    if teacher:
        assert abs(int(tc[0]) - int(ref[0])) <= near_tc + 2
    # shard invariance
    a = sim.run_lstm(h, L, [db], 250, 50, fast_math=2, teacher_forced=teacher, snr_idx=[2])[0]
    b = sim.run_lstm(h, L, [db], frames - 250, 300, fast_math=2, teacher_forced=teacher, snr_idx=[2])[0]
    assert ((a + b) == tc).all()


def test_decode_host_odd_batches_and_soft_outputs(sim):
    """nvd_mlp_decode_host (the reference's predictions.eval on host arrays): B = 1 / odd / multi-chunk with soft
    outputs on the bf16 path (sub-buffer alignment), equal to the device-buffer call."""
    from nvd import montecarlo as mc
    m = _model([14, 128, 64, 32, 128], [1, 1, 1, 1], seed=11, bias=0.1)
    h = sim.load_mlp(m)
    rng = np.random.default_rng(5)
    for B in (1, 3, 1001, (1 << 18) + 333):
        y = rng.standard_normal((B, 14)).astype(np.float32)
        pred, soft = sim.decode(h, y, "bf16", want_soft=True)
        y16 = mc.f32_to_bf16(y)
        soft_b, hard_b = sim.mlp_forward(h, sim.to_device(y16), "bf16", B, 128, "bf16", want_soft=True, hard_bytes=4)
        assert np.array_equal(soft, soft_b.download(np.float32, (B, 128)))
        assert np.array_equal(pred, hard_b.download(np.int32, (B,)))
        pred32, soft32 = sim.decode(h, y, "fp32", want_soft=True)
        assert np.abs(soft32 - omlp.forward(y, m.Ws, m.bs, m.acts)).max() <= 1e-5
