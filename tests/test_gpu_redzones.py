"""Red-zone tests: compute-sanitizer is closed on the GPU pool, so every kernel that writes a caller's device buffer is
run on buffers carved out of a larger allocation with 4 KiB canary zones on both sides, at awkward sizes (1, tile - 1,
tile + 1, ragged tails); any out-of-bounds WRITE of up to 4 KiB shows up as a damaged canary.  (Out-of-bounds reads
are bounded separately: inputs sit at the very end of their allocation minus the red zone, so a read far beyond would
fault; near misses are caught by the parity tests, whose results would change.)"""
import ctypes as C

import numpy as np
import pytest

from oracle import encoders as oenc, mlp as omlp

pytestmark = pytest.mark.gpu
ZONE = 4096


@pytest.fixture(scope="module")
def sim():
    from nvd import montecarlo
    with montecarlo.Simulator(device=0, seed=1) as s:
        yield s


class Guarded:
    """`nbytes` usable bytes at a 256-byte aligned device address with canary zones before and after."""

    def __init__(self, sim, nbytes, fill=0):
        from nvd import _lib
        self.sim, self.nbytes, self._lib = sim, int(nbytes), _lib
        self.pad = (-self.nbytes) % 256
        self.buf = sim.alloc(ZONE + self.nbytes + self.pad + ZONE)
        _lib.check(sim.lib.nvd_memset(sim.ctx, self.buf.ptr, 0xA5, self.buf.nbytes))
        self.ptr = self.buf.ptr + ZONE
        if self.nbytes:
            _lib.check(sim.lib.nvd_memset(sim.ctx, self.ptr, fill, self.nbytes))

    def upload(self, arr):
        arr = np.ascontiguousarray(arr)
        assert arr.nbytes == self.nbytes
        self._lib.check(self.sim.lib.nvd_memcpy_h2d(self.sim.ctx, self.ptr, arr.ctypes.data, arr.nbytes))
        self.sim.sync()
        return self

    def check(self, what=""):
        raw = self.buf.download(np.uint8, (self.buf.nbytes,))
        assert (raw[:ZONE] == 0xA5).all(), f"{what}: write BEFORE the buffer"
        assert (raw[ZONE + self.nbytes:] == 0xA5).all(), f"{what}: write BEYOND the buffer ({self.nbytes} bytes)"
        return raw[ZONE:ZONE + self.nbytes]


def _model(dims, acts, seed):
    from nvd import weights
    Ws, bs = omlp.init_fnn(dims, seed=seed, bias_scale=0.05)
    return weights.MlpModel(Ws, bs, acts)


@pytest.mark.parametrize("B", [1, 15, 17, 127, 129, 1000, 4099])
def test_source_encoder_channel_count_stay_inside_their_buffers(sim, B):
    from nvd import _lib
    lib, ctx = sim.lib, sim.ctx
    for L, code in ((7, 1), (7, 0), (100, 1), (518, 1), (13, 1)):
        S_in, S_out, n = (L + 7) // 8, (2 * L + 7) // 8, 2 * L
        info = Guarded(sim, B * S_in)
        _lib.check(lib.nvd_source(ctx, L, 12345, B, info.ptr))
        info.check(f"source L={L}")
        coded = Guarded(sim, B * S_out)
        _lib.check(lib.nvd_encode(ctx, code, info.ptr, L, B, coded.ptr))
        coded.check(f"encode L={L}")
        for dt, esz in ((_lib.F32, 4), (_lib.BF16, 2), (_lib.I32_TRUNC, 4), (_lib.BF16_TRUNC, 2)):
            y = Guarded(sim, B * n * esz)
            _lib.check(lib.nvd_channel(ctx, coded.ptr, n, B, C.c_float(0.7), 3, 99, dt, y.ptr))
            y.check(f"channel L={L} dtype={dt}")
        cnt = Guarded(sim, 32)
        _lib.check(lib.nvd_count_errors(ctx, info.ptr, info.ptr, S_in, L, B, 0, cnt.ptr))
        assert cnt.check("count").view(np.uint64).tolist() == [0, 0, B, B * L]
        if L == 518:
            pre = Guarded(sim, B * ((L + 6 + 7) // 8))
            _lib.check(lib.nvd_prepend_state(ctx, info.ptr, L, 33, 6, B, pre.ptr))
            pre.check("prepend")


@pytest.mark.parametrize("B", [1, 127, 129, 1000])
def test_decoders_stay_inside_their_buffers(sim, B):
    from nvd import _lib, montecarlo as mc
    lib, ctx = sim.lib, sim.ctx
    rng = np.random.default_rng(B)
    cases = [([14, 128, 64, 32, 128], [1, 1, 1, 1]),      # fused FNN kernel (staged form, TMA landing buffer)
             ([1036, 30, 2], [2, 2]),                     # K-streamed, cp.async rows
             ([1040, 30, 2], [2, 2]),                     # K-streamed, TMA rows
             ([200, 500, 500, 500, 500, 2], [1] * 5),     # wide-layer kernel
             ([30, 64, 10], [2, 0])]
    for dims, acts in cases:
        h = sim.load_mlp(_model(dims, acts, seed=len(dims)))
        y16 = mc.f32_to_bf16(rng.standard_normal((B, dims[0])).astype(np.float32))
        d_y = Guarded(sim, y16.nbytes).upload(y16)
        for prec in ("bf16", "fp32"):
            for hb in (1, 4):
                soft = Guarded(sim, B * dims[-1] * 4)
                hard = Guarded(sim, B * hb)
                _lib.check(lib.nvd_mlp_forward(ctx, h, d_y.ptr, _lib.BF16, B, _lib.PRECISIONS[prec], soft.ptr, hard.ptr, hb))
                sim.sync()
                s = soft.check(f"soft {dims} {prec}").view(np.float32).reshape(B, dims[-1])
                hd = hard.check(f"hard {dims} {prec}")
                idx = hd.view(np.int32) if hb == 4 else hd
                assert (idx == s.argmax(1)).all() or prec == "bf16"      # bf16 FNN path decides on the pre-activation
        d_y.check(f"input {dims}")
    # sliding-window decode: the last window ends exactly at the end of the stream
    hs = sim.load_mlp(_model([200, 500, 500, 500, 500, 2], [1] * 5, seed=5))
    n_sym = 2 * (B - 1) + 200
    stream = Guarded(sim, n_sym * 4).upload(rng.standard_normal(n_sym).astype(np.float32))
    for prec in ("bf16", "fp32"):
        bits = Guarded(sim, B)
        _lib.check(lib.nvd_window_decode(ctx, hs, stream.ptr, _lib.F32, n_sym, 2, B, _lib.PRECISIONS[prec], None, bits.ptr, 1))
        sim.sync()
        bits.check(f"window decode {prec}")


@pytest.mark.parametrize("B,T", [(1, 200), (129, 200), (300, 37)])
def test_recurrent_decoder_stays_inside_its_buffers(sim, B, T):
    from nvd import _lib, weights
    from oracle import lstm as olstm
    lib, ctx = sim.lib, sim.ctx
    p = olstm.init_params(20, seed=2)
    h = sim.load_lstm(weights.LstmModel(**p))
    rng = np.random.default_rng(T)
    y = Guarded(sim, B * T * 4).upload(rng.standard_normal((B, T)).astype(np.float32))
    dec = Guarded(sim, B * T * 4).upload(rng.standard_normal((B, T)).astype(np.float32))
    for fm in (0, 1, 2):
        for teacher in (True, False):
            logits = Guarded(sim, B * T * 4)
            hard = Guarded(sim, B * ((T + 7) // 8))
            _lib.check(lib.nvd_lstm_forward(ctx, h, y.ptr, _lib.F32, dec.ptr if teacher else None, T, T, B, fm, logits.ptr, hard.ptr))
            sim.sync()
            logits.check(f"lstm logits fm={fm}")
            hard.check(f"lstm hard fm={fm}")
    y.check("lstm input")
    info = Guarded(sim, B * 13).upload(rng.integers(0, 256, (B, 13)).astype(np.uint8))
    hardr = Guarded(sim, B * 25).upload(rng.integers(0, 256, (B, 25)).astype(np.uint8))
    cnt = Guarded(sim, 32)
    _lib.check(lib.nvd_count_row_errors(ctx, hardr.ptr, 25, info.ptr, 13, 100, B, cnt.ptr))
    c = cnt.check("count rows").view(np.uint64)
    assert c[2] == B and c[3] == 100 * B
    pred = Guarded(sim, B).upload(rng.integers(0, 2, B).astype(np.uint8))
    cnt2 = Guarded(sim, 32)
    _lib.check(lib.nvd_count_bit_errors(ctx, pred.ptr, info.ptr, 13, 99, B, cnt2.ptr))
    assert cnt2.check("count bit").view(np.uint64)[2] == B


def test_monte_carlo_steps_only_touch_their_counters(sim):
    from nvd import weights
    h = sim.load_mlp(weights.load_default_fnn())
    hb = sim.load_mlp(_model([1036, 30, 2], [2, 2], seed=9))
    from oracle import lstm as olstm
    hl = sim.load_lstm(weights.LstmModel(**olstm.init_params(20, seed=2)))
    n_snr = 3
    for step in ("fnn_fused", "fnn_staged", "bitwise_fused", "bitwise_staged", "lstm"):
        cnt = Guarded(sim, n_snr * 32)
        class Buf:      # the Simulator's run_*_async take anything with .ptr
            ptr = cnt.ptr
        if step == "fnn_fused":
            sim.run_async(h, "133171", 7, [0.0, 2.0, 4.0], 1000, Buf, 7, mode="fused")
        elif step == "fnn_staged":
            sim.run_async(h, "133171", 7, [0.0, 2.0, 4.0], 1000, Buf, 7, mode="staged", precision="bf16")
        elif step == "bitwise_fused":
            sim.run_bitwise_async(hb, 512, [0.0, 2.0, 4.0], 300, Buf, 7, mode="fused")
        elif step == "bitwise_staged":
            sim.run_bitwise_async(hb, 512, [0.0, 2.0, 4.0], 300, Buf, 7, mode="staged")
        else:
            sim.run_lstm_async(hl, 100, [0.0, 2.0, 4.0], 300, Buf, 7, fast_math=2)
        sim.sync()
        c = cnt.check(step).view(np.uint64).reshape(n_snr, 4)
        assert (c[:, 2] == (1000 if step.startswith("fnn") else 300)).all()
