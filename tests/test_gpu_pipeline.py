"""GPU tests of the whole Monte-Carlo step (encode -> channel -> decode -> count) through the C ABI:
fused == staged, shard invariance, oracle parity of the counters, host-buffer (end-to-end) form,
and size-independent properties at BASELINE.json's full size (2^24 frames / SNR point)."""
import os

import numpy as np
import pytest

from oracle import encoders as oenc, metrics as omet, mlp as omlp, pipeline as opipe

pytestmark = pytest.mark.gpu
SEED = 20170710


@pytest.fixture(scope="module")
def sim():
    from nvd import montecarlo
    with montecarlo.Simulator(device=0, seed=SEED) as s:
        yield s


@pytest.fixture(scope="module")
def model():
    from nvd import weights
    return weights.load_default_fnn()


@pytest.fixture(scope="module")
def handle(sim, model):
    return sim.load_mlp(model)


@pytest.mark.parametrize("frames,frame0", [(1, 0), (127, 5), (128, 0), (1000, 123), (65536 + 77, 1 << 33)])
def test_fused_equals_staged(sim, handle, frames, frame0):
    snr = [0.0, 3.0, 6.0]
    fused = sim.run(handle, "133171", 7, snr, frames, frame0, mode="fused")
    staged = sim.run(handle, "133171", 7, snr, frames, frame0, mode="staged", precision="bf16")
    assert (fused == staged).all()                      # bit-identical frames and arithmetic
    assert (fused[:, 2] == frames).all() and (fused[:, 3] == frames * 7).all()
    assert (fused[:, 1] <= frames).all() and (fused[:, 0] >= fused[:, 1]).all()


def test_shard_invariance(sim, handle):
    """Counters for N frames are identical for any split into rank ranges."""
    from nvd.montecarlo import shard_range
    N, snr = 100000, [1.0, 4.0]
    whole = sim.run(handle, "133171", 7, snr, N, 42, mode="fused")
    for world in (2, 3, 8):
        acc = np.zeros_like(whole)
        for rank in range(world):
            lo, n = shard_range(42, N, rank, world)
            acc += sim.run(handle, "133171", 7, snr, n, lo, mode="fused")
        assert (acc == whole).all()
    via_rank_api = sum(sim.run(handle, "133171", 7, snr, N, 42, mode="fused", rank=r, world=4) for r in range(4))
    assert (via_rank_api == whole).all()


@pytest.mark.parametrize("code,name", [(oenc.CODE_133171, "133171"), (oenc.CODE_57_REF, "57")])
def test_counters_match_oracle(sim, handle, model, code, name):
    """Same frames (Philox) through the CPU oracle: fp32 CUDA path must agree up to near-tie frames,
    bf16 tensor-core path up to the frames whose top-2 margin is inside the bf16 tolerance."""
    frames, snr_db = 1 << 15, 2.0
    ref, det = opipe.mc_step(model.Ws, model.bs, model.acts, 7, code, snr_db, 0, frames, seed=SEED, snr_idx=0,
                             noise="philox", precision="fp64", return_detail=True)
    margin = omlp.top2_margin(det["out"])
    fp32 = sim.run(handle, name, 7, [snr_db], frames, 0, mode="staged", precision="fp32", y_dtype="f32")[0]
    assert abs(int(fp32[1]) - int(ref[1])) <= int((margin < 1e-3).sum())
    assert abs(int(fp32[0]) - int(ref[0])) <= 7 * int((margin < 1e-3).sum())
    tc = sim.run(handle, name, 7, [snr_db], frames, 0, mode="fused")[0]
    assert abs(int(tc[1]) - int(ref[1])) <= int((margin < 5e-2).sum())
    assert tc[2] == ref[2] and tc[3] == ref[3]


def test_host_buffer_step_matches_device_step(sim, handle):
    frames = 50000
    info = oenc.class_to_packed(np.arange(frames) & 127, 7).astype(np.uint8)
    pred = np.zeros(frames, np.uint8)
    c_host = sim.run_host(handle, "133171", 7, 3.0, info, pred, frame0=0, precision="bf16", snr_idx=0)
    c_dev = sim.run(handle, "133171", 7, [3.0], frames, 0, mode="staged", precision="bf16")[0]
    assert (c_host == c_dev).all()
    errs = np.array([bin(int(a) ^ int(b)).count("1") for a, b in zip(pred[:5000], np.arange(5000) & 127)])
    assert errs.sum() <= c_host[0]


@pytest.mark.parametrize("frames", [1, 127, 50000, 148 * 4 * 128 * 24 + 4321])
def test_host_buffer_fused_step_equals_staged_step(sim, handle, frames):
    """nvd_mc_run_host on the fused kernel (info words from the host, sliced, copies overlapped) returns the same
    decisions and counters as the staged pipeline for ARBITRARY info words; last size spans two slices."""
    rng = np.random.default_rng(frames)
    info = rng.integers(0, 128, frames).astype(np.uint8)             # time-packed words, any order
    pred_f = np.full(frames, 255, np.uint8)
    pred_s = np.full(frames, 254, np.uint8)
    c_f = sim.run_host(handle, "133171", 7, 1.0, info, pred_f, frame0=777, snr_idx=3, mode="fused")
    c_s = sim.run_host(handle, "133171", 7, 1.0, info, pred_s, frame0=777, snr_idx=3, mode="staged")
    assert (c_f == c_s).all() and c_f[2] == frames and c_f[3] == 7 * frames
    assert (pred_f == pred_s).all()
    be, fe = omet.count_errors(pred_f, oenc.packed_to_class(info, 7), 7)
    assert [be, fe] == [int(c_f[0]), int(c_f[1])]
    c_nopred = sim.run_host(handle, "133171", 7, 1.0, info, None, frame0=777, snr_idx=3, mode="fused")
    assert (c_nopred == c_f).all()
    # pinned buffers take the streaming form: one launch, info words arriving while the kernel runs (flagged slices),
    # decisions stored straight into the mapped host buffer; repeated to catch stale flags / cached info bytes
    info_p, pred_p = sim.pinned(frames), sim.pinned(frames)
    os.environ["NVD_HOST_STEP"] = "streaming"        # opt-in; the sliced form is the default
    for rep in range(3):
        info_r = np.roll(info, rep)
        info_p[:] = info_r
        pred_p[:] = 253
        c_p = sim.run_host(handle, "133171", 7, 1.0, info_p, pred_p, frame0=777, snr_idx=3, mode="fused")
        if rep == 0:
            assert (c_p == c_s).all() and (pred_p == pred_s).all()
        else:
            pred_r = np.full(frames, 252, np.uint8)
            c_r = sim.run_host(handle, "133171", 7, 1.0, info_r, pred_r, frame0=777, snr_idx=3, mode="staged")
            assert (c_p == c_r).all() and (pred_p == pred_r).all()
    os.environ.pop("NVD_HOST_STEP")
    c_p2 = sim.run_host(handle, "133171", 7, 1.0, info_p, None, frame0=777, snr_idx=3, mode="fused")
    assert (c_p2 == c_p).all()


def test_full_size_properties(sim, handle):
    """BASELINE.json configs[1] size: 2^24 frames per SNR point.  Size-independent checks: totals,
    monotone BER in SNR, fused == sum of two half ranges, noiseless channel decodes the code book."""
    N = 1 << 24
    snr = [0.0, 4.0, 8.0]
    c = sim.run(handle, "133171", 7, snr, N, 0, mode="fused")
    assert (c[:, 2] == N).all() and (c[:, 3] == 7 * N).all()
    ber = c[:, 0] / c[:, 3]
    assert ber[0] > ber[1] > ber[2]
    halves = sim.run(handle, "133171", 7, snr, N // 2, 0, mode="fused") + sim.run(handle, "133171", 7, snr, N // 2, N // 2, mode="fused")
    assert (halves == c).all()


def test_bad_arguments_fail_loudly(sim, handle):
    from nvd import NvdError
    with pytest.raises(NvdError):
        sim.run(handle, "133171", 9, [1.0], 10)              # L does not match the model
    with pytest.raises(NvdError):
        sim.run(99, "133171", 7, [1.0], 10)                  # bad handle
    with pytest.raises(NvdError):
        sim.run(handle, "133171", 7, [1.0], 10, mode="fused", precision="fp32")


# ------------------------------------------------------------------ BER-curve parity (north-star criterion)
def _sweep_against_oracle(sim, model, code_name, code, snrs, frames):
    from nvd import montecarlo as mc
    h = sim.load_mlp(model)
    fp32 = sim.run(h, code_name, 7, snrs, frames, 0, mode="staged", precision="fp32", y_dtype="f32")
    bf16 = sim.run(h, code_name, 7, snrs, frames, 0, mode="fused")
    prev = None
    for i, db in enumerate(snrs):
        ref, det = opipe.mc_step(model.Ws, model.bs, model.acts, 7, code, float(db), 0, frames, seed=SEED, snr_idx=i,
                                 noise="philox", precision="fp64", return_detail=True)
        margin = omlp.top2_margin(det["out"])
        near = int((margin < 1e-3).sum())
        # fp32 path: the same decisions except on near-tie frames (|margin| < 1e-3)
        assert abs(int(fp32[i][1]) - int(ref[1])) <= near and abs(int(fp32[i][0]) - int(ref[0])) <= 7 * near, (db, fp32[i], ref, near)
        # bf16 tensor-core path: BER inside the oracle's 95 % Monte-Carlo (Wilson) interval on the same frames
        lo, hi = omet.wilson_interval(int(ref[0]), int(ref[3]))
        ber = float(bf16[i][0]) / float(bf16[i][3])
        assert lo <= ber <= hi, (db, ber, lo, hi)
        flo, fhi = omet.wilson_interval(int(ref[1]), int(ref[2]))
        assert flo <= float(bf16[i][1]) / float(bf16[i][2]) <= fhi
        if prev is not None and ref[0] > 50:
            assert ref[0] < prev                      # BER falls with SNR
        prev = ref[0]
    return bf16


def test_ber_sweep_133171_matches_oracle(sim, model):
    """BASELINE.json configs[1] curve: (133,171) FNN, 0..8 dB, 2^18 frames per point, GPU vs CPU oracle on the SAME
    Philox frames — fp32 counters equal up to near ties, bf16 BER/FER inside the oracle's 95 % interval everywhere."""
    bf16 = _sweep_against_oracle(sim, model, "133171", oenc.CODE_133171, list(range(0, 9)), 1 << 18)
    ber = bf16[:, 0] / bf16[:, 3]
    assert ber[0] > 1e-2 and ber[4] < 1e-3            # a trained decoder (weights fixture), not a random one


def test_ber_sweep_57_matches_oracle(sim):
    """BASELINE.json configs[0]: the reference's (7,5) encoder (quirk included) + FNN decoder, 0..6 dB."""
    from nvd import weights
    m57 = weights.load_fnn_57()
    bf16 = _sweep_against_oracle(sim, m57, "57", oenc.CODE_57_REF, list(range(0, 7)), 1 << 18)
    # the reference's own test-set size (128 classes x 10 000, src/fnn_viterbi.py:58): fused == staged at that size
    h = sim.load_mlp(m57)
    n = 1_280_000
    a = sim.run(h, "57", 7, [0.0, 3.0, 6.0], n, 0, mode="fused")
    b = sim.run(h, "57", 7, [0.0, 3.0, 6.0], n, 0, mode="staged", precision="bf16")
    assert (a == b).all() and a[0, 0] > a[1, 0] > a[2, 0]
    assert bf16[0, 0] / bf16[0, 3] < 0.2


@pytest.mark.gpu
def test_peer_memory_exchange_single_rank():
    """csrc/comm.cu peer-memory all-reduce with world = 1 (the kernels run against the rank's own exchange block):
    identity on the values, no carry-over between exchanges (the two alternating areas are cleared), any n up to 448,
    and the Monte-Carlo counters pass through it unchanged."""
    from nvd import montecarlo
    with montecarlo.Simulator(0) as s:
        assert s.comm_p2p_open(0, 1, all_gather=lambda b, n: bytes(b))
        assert s.comm_info()["transport"] == "nvlink-peer-memory"
        rng = np.random.default_rng(11)
        for rnd in range(9):
            n = [4, 448, 1, 36, 100][rnd % 5]
            mine = rng.integers(0, 1 << 62, n, dtype=np.uint64)
            buf = s.to_device(mine)
            s.allreduce_counters(buf, n)
            assert np.array_equal(buf.download(np.uint64, (n,)), mine), (rnd, n)
            s.free(buf)
        big = s.to_device(np.arange(449, dtype=np.uint64))        # above the block's capacity: the NCCL path, a no-op without a communicator
        s.allreduce_counters(big, 449)
        assert np.array_equal(big.download(np.uint64, (449,)), np.arange(449, dtype=np.uint64))
