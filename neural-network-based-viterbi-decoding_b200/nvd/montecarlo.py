"""Simulator: the Python face of libnvd's context — device buffers, the five kernels and the
Monte-Carlo step, staged or fused, single GPU or sharded over torch.distributed ranks.

Sharding: rank g of G takes the contiguous global frame range
[frame0 + g*N/G, frame0 + (g+1)*N/G) of every SNR point.  Info word and noise are pure functions
of the GLOBAL frame index (counter-based Philox), so the counters are independent of G; the only
exchange is one all-reduce(sum) of the [n_snr, 4] uint64 counters.
"""
from __future__ import annotations

import os

import ctypes as C
import math

import numpy as np

from . import _lib
from ._lib import NvdError, check
from .weights import LstmModel, MlpModel

DEFAULT_SEED = 20170710


def snr_db_to_sigma(snr_db):
    """sigma = sqrt(1 / (2 * 10^(dB/10))) — /root/reference/src/fnn_viterbi.py:89-92."""
    return np.sqrt(0.5 / (10.0 ** (np.asarray(snr_db, dtype=np.float64) / 10.0)))


def bf16_to_f32(u16: np.ndarray) -> np.ndarray:
    return (np.asarray(u16, dtype=np.uint16).astype(np.uint32) << 16).view(np.float32)


def f32_to_bf16(x: np.ndarray) -> np.ndarray:
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    return ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)


def class_to_packed(cls, L: int) -> np.ndarray:
    """MSB-first class index (generate_class_bits row order, /root/reference/src/fnn_viterbi.py:283)
    -> time-packed info word (bit t = info bit at time t): bit reversal over L bits."""
    cls = np.asarray(cls).astype(np.int64)
    out = np.zeros_like(cls)
    for t in range(L):
        out |= ((cls >> (L - 1 - t)) & 1) << t
    return out


def shard_range(frame0: int, frames: int, rank: int, world: int):
    """Contiguous split of [frame0, frame0+frames) — identical totals for any world size."""
    lo = frame0 + (frames * rank) // world
    hi = frame0 + (frames * (rank + 1)) // world
    return lo, hi - lo


class DeviceBuffer:
    def __init__(self, sim: "Simulator", nbytes: int):
        self.sim, self.nbytes = sim, int(nbytes)
        p = C.c_void_p()
        check(sim.lib.nvd_alloc(sim.ctx, self.nbytes, C.byref(p)))
        self.ptr = p.value or 0

    def free(self):
        if self.ptr and self.sim.ctx:
            check(self.sim.lib.nvd_free(self.sim.ctx, self.ptr))
        self.ptr = 0

    def upload(self, arr: np.ndarray):
        arr = np.ascontiguousarray(arr)
        assert arr.nbytes <= self.nbytes
        check(self.sim.lib.nvd_memcpy_h2d(self.sim.ctx, self.ptr, arr.ctypes.data, arr.nbytes))
        self.sim.sync()  # pageable source: keep it alive until the copy is done
        return self

    def download(self, dtype, shape) -> np.ndarray:
        out = np.empty(shape, dtype=dtype)
        assert out.nbytes <= self.nbytes
        check(self.sim.lib.nvd_memcpy_d2h(self.sim.ctx, out.ctypes.data, self.ptr, out.nbytes))
        self.sim.sync()
        return out


class Simulator:
    """One libnvd context on one GPU.  Use as a context manager."""

    def __init__(self, device: int = 0, seed: int = DEFAULT_SEED):
        self.lib = _lib.load()
        ctx = C.c_void_p()
        check(self.lib.nvd_ctx_create(device, seed, C.byref(ctx)))
        self.ctx = ctx
        self.device = device
        self.seed = seed
        self._bufs = []
        self._n_out = {}
        self.rank, self.world = 0, 1
        self.p2p = False

    # -- lifetime ------------------------------------------------------------------------------
    def close(self):
        if self.ctx:
            for b in self._bufs:
                try:
                    b.free()
                except NvdError:
                    pass
            self._bufs = []
            self.lib.nvd_comm_destroy(self.ctx)
            self.lib.nvd_ctx_destroy(self.ctx)
            self.ctx = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        check(self.lib.nvd_sync(self.ctx))

    def set_seed(self, seed: int):
        check(self.lib.nvd_set_seed(self.ctx, seed))
        self.seed = seed

    def device_info(self):
        sm, ma, mi, mem = C.c_int(), C.c_int(), C.c_int(), C.c_size_t()
        check(self.lib.nvd_device_info(self.ctx, C.byref(sm), C.byref(ma), C.byref(mi), C.byref(mem)))
        return dict(sm_count=sm.value, cc=(ma.value, mi.value), total_mem=mem.value)

    # -- multi-GPU exchange step (include/nvd.h: nvd_comm_*) -------------------------------------------
    def comm_init(self, rank: int, world: int, broadcast=None):
        """Create this context's NCCL communicator.  `broadcast(bytes_or_None, nbytes) -> bytes` ships rank 0's
        unique id to the other ranks (default: torch.distributed, nvd.dist.broadcast_bytes)."""
        if world <= 1:
            return
        if broadcast is None:
            from . import dist
            broadcast = dist.broadcast_bytes
        raw = None
        if rank == 0:
            buf = C.create_string_buffer(_lib.COMM_ID_BYTES)
            check(self.lib.nvd_comm_unique_id(buf))
            raw = buf.raw
        raw = broadcast(raw, _lib.COMM_ID_BYTES)
        check(self.lib.nvd_comm_init(self.ctx, rank, world, C.create_string_buffer(raw, _lib.COMM_ID_BYTES)))
        self.rank, self.world = rank, world
        if os.environ.get("NVD_P2P", "1") != "0":
            self.comm_p2p_open(rank, world)

    def comm_p2p_open(self, rank: int, world: int, all_gather=None) -> bool:
        """Map every rank's exchange block (CUDA IPC) so that nvd_comm_allreduce_u64 goes over NVLink peer memory
        instead of NCCL.  `all_gather(bytes, nbytes) -> bytes` concatenates the ranks' handles in rank order (default:
        torch.distributed).  Returns False — and the NCCL transport stays in use — when a peer cannot be mapped; the
        decision is made collectively, so that all ranks use the same transport."""
        if all_gather is None:
            from . import dist
            all_gather = dist.all_gather_bytes
        buf = C.create_string_buffer(_lib.COMM_IPC_BYTES)
        check(self.lib.nvd_comm_p2p_handle(self.ctx, buf))
        handles = all_gather(buf.raw, _lib.COMM_IPC_BYTES)
        if len(handles) != world * _lib.COMM_IPC_BYTES:      # no channel that reaches every rank (e.g. torch.distributed not
            self.p2p = False                                 # initialised and no all_gather given): stay on NCCL
            check(self.lib.nvd_comm_p2p_close(self.ctx))
            return False
        status = self.lib.nvd_comm_p2p_open(self.ctx, rank, world, C.create_string_buffer(handles, len(handles)))
        oks = all_gather(bytes([1 if status == 0 else 0]), 1)
        self.p2p = all(b == 1 for b in oks)
        if status == 0 and not self.p2p:            # somebody else failed: everybody goes back to NCCL
            check(self.lib.nvd_comm_p2p_close(self.ctx))
        return self.p2p

    def comm_info(self):
        r, w, v = C.c_int(), C.c_int(), C.c_int()
        check(self.lib.nvd_comm_info(self.ctx, C.byref(r), C.byref(w), C.byref(v)))
        return dict(rank=r.value, world=w.value, nccl_version=v.value, transport="nvlink-peer-memory" if self.p2p else "nccl")

    def allreduce_counters(self, counters: "DeviceBuffer", n_u64: int):
        """In-place sum over all ranks of n uint64 device values on the library's stream (asynchronous): over NVLink
        peer memory when comm_p2p_open succeeded, else ncclAllReduce."""
        check(self.lib.nvd_comm_allreduce_u64(self.ctx, counters.ptr, n_u64))

    # -- memory / timing -------------------------------------------------------------------------
    def alloc(self, nbytes: int) -> DeviceBuffer:
        b = DeviceBuffer(self, max(int(nbytes), 16))
        self._bufs.append(b)
        return b

    def free(self, buf: DeviceBuffer):
        buf.free()
        if buf in self._bufs:
            self._bufs.remove(buf)

    def to_device(self, arr: np.ndarray) -> DeviceBuffer:
        arr = np.ascontiguousarray(arr)
        return self.alloc(arr.nbytes).upload(arr)

    def zeros(self, nbytes: int) -> DeviceBuffer:
        b = self.alloc(nbytes)
        check(self.lib.nvd_memset(self.ctx, b.ptr, 0, b.nbytes))
        return b

    def pinned(self, nbytes: int, dtype=np.uint8):
        """Pinned host array (numpy view over nvd_host_alloc memory); freed with the process."""
        p = C.c_void_p()
        check(self.lib.nvd_host_alloc(self.ctx, int(nbytes), C.byref(p)))
        raw = (C.c_uint8 * int(nbytes)).from_address(p.value)
        return np.frombuffer(raw, dtype=dtype)

    def event(self):
        e = C.c_void_p()
        check(self.lib.nvd_event_create(self.ctx, C.byref(e)))
        return e

    def record(self, ev):
        check(self.lib.nvd_event_record(self.ctx, ev))

    def elapsed_ms(self, e0, e1) -> float:
        ms = C.c_float()
        check(self.lib.nvd_event_elapsed_ms(self.ctx, e0, e1, C.byref(ms)))
        return ms.value

    # -- models ----------------------------------------------------------------------------------
    def load_mlp(self, model: MlpModel) -> int:
        n = len(model.Ws)
        dims = (C.c_int * (n + 1))(*model.dims)
        acts = (C.c_int * n)(*[int(a) for a in model.acts])
        Ws = [np.ascontiguousarray(W, dtype=np.float32) for W in model.Ws]
        bs = [np.ascontiguousarray(b, dtype=np.float32) for b in model.bs]
        Wp = (C.c_void_p * n)(*[W.ctypes.data for W in Ws])
        bp = (C.c_void_p * n)(*[b.ctypes.data for b in bs])
        h = C.c_int(-1)
        check(self.lib.nvd_mlp_load(self.ctx, n, dims, acts, Wp, bp, C.byref(h)))
        self._n_out[h.value] = model.dims[-1]
        return h.value

    def mlp_tc_supported(self, h: int) -> bool:
        s = C.c_int(0)
        check(self.lib.nvd_mlp_tc_supported(self.ctx, h, C.byref(s)))
        return bool(s.value)

    def load_lstm(self, m: LstmModel) -> int:
        arrs = [np.ascontiguousarray(a, dtype=np.float32) for a in (m.enc_W, m.enc_b, m.dec_W, m.dec_b, m.lin_w, m.lin_b)]
        h = C.c_int(-1)
        check(self.lib.nvd_lstm_load(self.ctx, m.H, *[a.ctypes.data for a in arrs], C.byref(h)))
        return h.value

    # -- kernels on device buffers -------------------------------------------------------------------
    def source(self, L: int, frame0: int, B: int) -> DeviceBuffer:
        out = self.alloc(B * ((L + 7) // 8))
        check(self.lib.nvd_source(self.ctx, L, frame0, B, out.ptr))
        return out

    def encode(self, code, info: DeviceBuffer, L: int, B: int) -> DeviceBuffer:
        out = self.alloc(B * ((2 * L + 7) // 8))
        check(self.lib.nvd_encode(self.ctx, _code(code), info.ptr, L, B, out.ptr))
        return out

    def channel(self, coded: DeviceBuffer, n: int, B: int, sigma: float, snr_idx: int = 0, frame0: int = 0,
                out_dtype="f32") -> DeviceBuffer:
        dt = _lib.DTYPES[out_dtype]
        out = self.alloc(B * n * (2 if dt == _lib.BF16 else 4))
        check(self.lib.nvd_channel(self.ctx, coded.ptr, n, B, float(sigma), snr_idx, frame0, dt, out.ptr))
        return out

    def mlp_forward(self, h: int, y: DeviceBuffer, y_dtype, B: int, n_out: int, precision="fp32", want_soft=True,
                    hard_bytes=4):
        soft = self.alloc(B * n_out * 4) if want_soft else None
        hard = self.alloc(B * hard_bytes) if hard_bytes else None
        check(self.lib.nvd_mlp_forward(self.ctx, h, y.ptr, _lib.DTYPES[y_dtype], B, _lib.PRECISIONS[precision],
                                       soft.ptr if soft else None, hard.ptr if hard else None, hard_bytes or 4))
        return soft, hard

    def count_errors(self, pred: DeviceBuffer, label: DeviceBuffer, stride: int, L: int, B: int, flags: int,
                     counters: DeviceBuffer):
        check(self.lib.nvd_count_errors(self.ctx, pred.ptr, label.ptr, stride, L, B, flags, counters.ptr))

    def window_decode(self, h: int, stream: DeviceBuffer, stream_dtype, n_symbols: int, stride: int, n_windows: int,
                      n_out: int, precision="fp32", want_soft=True, hard_bytes=4):
        """Sliding-window decode of one long symbol stream in device memory (nvd_window_decode)."""
        soft = self.alloc(n_windows * n_out * 4) if want_soft else None
        hard = self.alloc(n_windows * hard_bytes)
        check(self.lib.nvd_window_decode(self.ctx, h, stream.ptr, _lib.DTYPES[stream_dtype], n_symbols, stride, n_windows,
                                         _lib.PRECISIONS[precision], soft.ptr if soft else None, hard.ptr, hard_bytes))
        return soft, hard

    def lstm_forward(self, h: int, y: DeviceBuffer, y_dtype, dec_in, T_enc: int, T_dec: int, B: int, fast_math=False,
                     want_logits=True, want_hard=True):
        logits = self.alloc(B * T_dec * 4) if want_logits else None
        hard = self.alloc(B * ((T_dec + 7) // 8)) if want_hard else None
        check(self.lib.nvd_lstm_forward(self.ctx, h, y.ptr, _lib.DTYPES[y_dtype], dec_in.ptr if dec_in else None,
                                        T_enc, T_dec, B, int(fast_math), logits.ptr if logits else None,
                                        hard.ptr if hard else None))
        return logits, hard

    # -- numpy-level conveniences (what the drop-in modules call) --------------------------------------
    def encode_bits(self, bits: np.ndarray, code="133171") -> np.ndarray:
        """{0,1} [B, L] (or [L]) -> coded {0,1} int32 [B, 2L] through the K1 kernel."""
        bits = np.asarray(bits)
        single = bits.ndim == 1
        b2 = (np.mod(bits.reshape(-1, bits.shape[-1]), 2)).astype(np.uint8)
        B, L = b2.shape
        packed = np.packbits(b2, axis=1, bitorder="little")
        d_in = self.to_device(packed)
        d_out = self.encode(code, d_in, L, B)
        out = d_out.download(np.uint8, (B, (2 * L + 7) // 8))
        self.free(d_in)
        self.free(d_out)
        coded = np.unpackbits(out, axis=1, bitorder="little")[:, : 2 * L].astype(np.int32)
        return coded[0] if single else coded

    def channel_bits(self, coded: np.ndarray, sigma: float, snr_idx=0, frame0=0, out_dtype="f32") -> np.ndarray:
        coded = np.asarray(coded)
        single = coded.ndim == 1
        c2 = coded.reshape(-1, coded.shape[-1]).astype(np.uint8)
        B, n = c2.shape
        d_in = self.to_device(np.packbits(c2, axis=1, bitorder="little"))
        d_y = self.channel(d_in, n, B, sigma, snr_idx, frame0, out_dtype)
        if out_dtype == "bf16":
            y = bf16_to_f32(d_y.download(np.uint16, (B, n)))
        elif out_dtype == "i32_trunc":
            y = d_y.download(np.int32, (B, n))
        else:
            y = d_y.download(np.float32, (B, n))
        self.free(d_in)
        self.free(d_y)
        return y[0] if single else y

    def decode(self, h: int, y: np.ndarray, precision="fp32", want_soft=True):
        """Noisy blocks [B, n_in] fp32 on the host -> (predictions int32 [B], soft [B, n_out] or None).
        This is `predictions_tensor.eval(feed_dict={x: test_set})` of the reference."""
        y = np.ascontiguousarray(y, dtype=np.float32)
        B = y.shape[0]
        n_out = self._n_out[h]
        soft = np.empty((B, n_out), np.float32) if want_soft else None
        hard = np.empty(B, np.int32)
        check(self.lib.nvd_mlp_decode_host(self.ctx, h, y.ctypes.data, B, _lib.PRECISIONS[precision],
                                           soft.ctypes.data if want_soft else None, hard.ctypes.data))
        return hard, soft

    # -- Monte-Carlo step --------------------------------------------------------------------------------
    def _mc_config(self, h, code, L, precision, mode, y_dtype, frame0, frames):
        cfg = _lib.McConfig()
        cfg.struct_bytes = C.sizeof(_lib.McConfig)
        cfg.model, cfg.code, cfg.L = h, _code(code), L
        cfg.precision = _lib.PRECISIONS[precision]
        cfg.mode = _lib.MC_FUSED if mode == "fused" else _lib.MC_STAGED
        cfg.y_dtype = _lib.DTYPES[y_dtype]
        cfg.frame0, cfg.frames = int(frame0), int(frames)
        return cfg

    def run_async(self, h, code, L, snr_db, frames, counters: DeviceBuffer, frame0=0, mode="fused", precision="bf16",
                  y_dtype=None, snr_idx=None):
        """Enqueue one Monte-Carlo pass over [frame0, frame0+frames) at every SNR point;
        counters ([n_snr, 4] uint64 on the device) are accumulated into."""
        snr_db = np.atleast_1d(np.asarray(snr_db, dtype=np.float32))
        if y_dtype is None:
            y_dtype = "bf16" if precision == "bf16" else "f32"
        idx = np.arange(len(snr_db), dtype=np.int32) if snr_idx is None else np.asarray(snr_idx, dtype=np.int32)
        cfg = self._mc_config(h, code, L, precision, mode, y_dtype, frame0, frames)
        check(self.lib.nvd_mc_run(self.ctx, C.byref(cfg), snr_db.ctypes.data_as(C.POINTER(C.c_float)),
                                  idx.ctypes.data_as(C.POINTER(C.c_int32)), len(snr_db), counters.ptr))

    def run(self, h, code, L, snr_db, frames, frame0=0, mode="fused", precision="bf16", y_dtype=None, snr_idx=None,
            rank=0, world=1, all_reduce=None) -> np.ndarray:
        """Synchronous Monte-Carlo pass -> counters uint64 [n_snr, 4] = {bit errors, frame errors,
        frames, bits}.  With world > 1 this rank runs its shard and `all_reduce(np.int64 array)`
        (e.g. a torch.distributed NCCL all-reduce) sums the counters."""
        snr_db = np.atleast_1d(np.asarray(snr_db, dtype=np.float32))
        lo, n = shard_range(frame0, frames, rank, world)
        counters = self.zeros(len(snr_db) * 32)
        self.run_async(h, code, L, snr_db, n, counters, lo, mode, precision, y_dtype, snr_idx)
        out = counters.download(np.uint64, (len(snr_db), 4))
        self.free(counters)
        if world > 1 and all_reduce is not None:
            out = all_reduce(out.astype(np.int64)).astype(np.uint64)
        return out

    def run_host(self, h, code, L, snr_db: float, info_host: np.ndarray, pred_host=None, frame0=0, precision="bf16",
                 y_dtype=None, snr_idx=0, mode="staged") -> np.ndarray:
        """End-to-end step with HOST buffers: info words in, decisions + counters out (H2D/D2H inside).
        mode "fused": the fused tcgen05 kernel over slices, copies overlapped on two copy streams."""
        if y_dtype is None:
            y_dtype = "bf16" if precision == "bf16" else "f32"
        B = info_host.shape[0]
        cfg = self._mc_config(h, code, L, precision, mode, y_dtype, frame0, B)
        counters = np.zeros(4, dtype=np.uint64)
        check(self.lib.nvd_mc_run_host(self.ctx, C.byref(cfg), float(snr_db), int(snr_idx), info_host.ctypes.data,
                                       pred_host.ctypes.data if pred_host is not None else None, counters.ctypes.data))
        return counters

    def run_bitwise_async(self, h, L, snr_db, frames, counters: DeviceBuffer, frame0=0, mode="fused", precision="bf16",
                          code="133171", n_state_bits=6, state=0, bit_position=0, trunc=True, snr_idx=None):
        """Enqueue the per-bit predictor's Monte-Carlo step (nvd_mc_run_bitwise): Philox info bits, start-state prefix,
        encoder, channel (int32-truncated like the reference's data sets when `trunc`), 1036-30-2 forward, decision
        against info bit `bit_position`; counters [n_snr, 4] accumulated on the device."""
        snr_db = np.atleast_1d(np.asarray(snr_db, dtype=np.float32))
        idx = np.arange(len(snr_db), dtype=np.int32) if snr_idx is None else np.asarray(snr_idx, dtype=np.int32)
        cfg = _lib.McBitwiseConfig()
        cfg.struct_bytes = C.sizeof(_lib.McBitwiseConfig)
        cfg.model, cfg.code, cfg.L = h, _code(code), L
        cfg.n_state_bits, cfg.state, cfg.bit_position = n_state_bits, state, bit_position
        cfg.precision = _lib.PRECISIONS[precision]
        cfg.mode = _lib.MC_FUSED if mode == "fused" else _lib.MC_STAGED
        cfg.trunc = int(bool(trunc))
        cfg.frame0, cfg.frames = int(frame0), int(frames)
        check(self.lib.nvd_mc_run_bitwise(self.ctx, C.byref(cfg), snr_db.ctypes.data_as(C.POINTER(C.c_float)),
                                          idx.ctypes.data_as(C.POINTER(C.c_int32)), len(snr_db), counters.ptr))

    def run_bitwise(self, h, L, snr_db, frames, frame0=0, rank=0, world=1, **kw) -> np.ndarray:
        """Synchronous form; with world > 1 this rank runs its shard and the library all-reduces the counters."""
        snr_db = np.atleast_1d(np.asarray(snr_db, dtype=np.float32))
        lo, n = shard_range(frame0, frames, rank, world)
        counters = self.zeros(len(snr_db) * 32)
        self.run_bitwise_async(h, L, snr_db, n, counters, lo, **kw)
        if world > 1:
            self.allreduce_counters(counters, 4 * len(snr_db))
        out = counters.download(np.uint64, (len(snr_db), 4))
        self.free(counters)
        return out

    def run_lstm_async(self, h, L, snr_db, frames, counters: DeviceBuffer, frame0=0, fast_math=2, teacher_forced=True,
                       code="133171", trunc=True, snr_idx=None):
        """Enqueue the recurrent decoder's Monte-Carlo step (nvd_mc_run_lstm): Philox info bits, encoder, channel,
        LSTM seq2seq over 2L + 2L steps, decision bits of the first L decoder steps against the info bits."""
        snr_db = np.atleast_1d(np.asarray(snr_db, dtype=np.float32))
        idx = np.arange(len(snr_db), dtype=np.int32) if snr_idx is None else np.asarray(snr_idx, dtype=np.int32)
        cfg = _lib.McLstmConfig()
        cfg.struct_bytes = C.sizeof(_lib.McLstmConfig)
        cfg.model, cfg.code, cfg.L = h, _code(code), L
        cfg.fast_math, cfg.teacher_forced, cfg.trunc = int(fast_math), int(bool(teacher_forced)), int(bool(trunc))
        cfg.frame0, cfg.frames = int(frame0), int(frames)
        check(self.lib.nvd_mc_run_lstm(self.ctx, C.byref(cfg), snr_db.ctypes.data_as(C.POINTER(C.c_float)),
                                       idx.ctypes.data_as(C.POINTER(C.c_int32)), len(snr_db), counters.ptr))

    def run_lstm(self, h, L, snr_db, frames, frame0=0, rank=0, world=1, **kw) -> np.ndarray:
        snr_db = np.atleast_1d(np.asarray(snr_db, dtype=np.float32))
        lo, n = shard_range(frame0, frames, rank, world)
        counters = self.zeros(len(snr_db) * 32)
        self.run_lstm_async(h, L, snr_db, n, counters, lo, **kw)
        if world > 1:
            self.allreduce_counters(counters, 4 * len(snr_db))
        out = counters.download(np.uint64, (len(snr_db), 4))
        self.free(counters)
        return out

    def count_bit_errors(self, pred: np.ndarray, info_bits: np.ndarray, bit_position: int) -> np.ndarray:
        """Decisions (0/1) [B] against bit `bit_position` of info rows {0,1} [B, L] on K5 -> counters uint64[4]."""
        pred = np.ascontiguousarray(pred, dtype=np.uint8)
        packed = np.packbits(np.asarray(info_bits, dtype=np.uint8) & 1, axis=1, bitorder="little")
        d_p, d_i, c = self.to_device(pred), self.to_device(packed), self.zeros(32)
        check(self.lib.nvd_count_bit_errors(self.ctx, d_p.ptr, d_i.ptr, packed.shape[1], bit_position, len(pred), c.ptr))
        out = c.download(np.uint64, (4,))
        for b in (d_p, d_i, c):
            self.free(b)
        return out

    def count_row_errors(self, pred_bits: np.ndarray, info_bits: np.ndarray) -> np.ndarray:
        """Decision rows {0,1} [B, >= L] against info rows {0,1} [B, L] on K5 (first L bits) -> counters uint64[4]."""
        info_bits = np.asarray(info_bits, dtype=np.uint8) & 1
        L = info_bits.shape[1]
        pp = np.packbits(np.asarray(pred_bits, dtype=np.uint8) & 1, axis=1, bitorder="little")
        ip = np.packbits(info_bits, axis=1, bitorder="little")
        d_p, d_i, c = self.to_device(pp), self.to_device(ip), self.zeros(32)
        check(self.lib.nvd_count_row_errors(self.ctx, d_p.ptr, pp.shape[1], d_i.ptr, ip.shape[1], L, len(pp), c.ptr))
        out = c.download(np.uint64, (4,))
        for b in (d_p, d_i, c):
            self.free(b)
        return out

    def host_step(self, h, code, L, snr_db: float, info_host: np.ndarray, pred_host=None, frame0=0, precision="bf16",
                  snr_idx=0, mode="fused"):
        """A prepared end-to-end step: the configuration struct and the ctypes arguments are built ONCE; calling the
        returned function is nothing but the C-ABI call nvd_mc_run_host (counters come back in the returned array,
        overwritten by every call)."""
        cfg = self._mc_config(h, code, L, precision, mode, "bf16" if precision == "bf16" else "f32", frame0, info_host.shape[0])
        counters = np.zeros(4, dtype=np.uint64)
        args = (self.ctx, C.byref(cfg), C.c_float(float(snr_db)), C.c_int(int(snr_idx)), C.c_void_p(info_host.ctypes.data),
                C.c_void_p(pred_host.ctypes.data) if pred_host is not None else None, C.c_void_p(counters.ctypes.data))
        fn = self.lib.nvd_mc_run_host

        def step(_keep=(cfg, info_host, pred_host)):
            check(fn(*args))
            return counters
        return step

    def stage_ms(self):
        ms = (C.c_float * 5)()
        check(self.lib.nvd_mc_stage_ms(self.ctx, ms))
        return dict(zip(("source", "encode", "channel", "decode", "count"), list(ms)))

    def set_stage_timing(self, on: bool):
        check(self.lib.nvd_mc_set_stage_timing(self.ctx, int(on)))

    def sweep(self, h, code, L, snr_db, min_frame_errors=100, max_frames=1 << 34, chunk_frames=1 << 26, mode="fused",
              precision="bf16", rank=0, world=1, all_reduce=None, frame0=0):
        """Deep-BER driver: per SNR point, add chunks of frames until >= min_frame_errors frame
        errors (global, after the all-reduce) or max_frames.  Returns counters [n_snr, 4].
        Without an `all_reduce` callable this is ONE C-ABI call (nvd_mc_sweep): the library shards every chunk over
        the ranks of its communicator (comm_init) and all-reduces the counters on its own stream."""
        snr_db = np.atleast_1d(np.asarray(snr_db, dtype=np.float32))
        if all_reduce is None and (world == 1 or self.world == world):
            cfg = self._mc_config(h, code, L, precision, mode, "bf16" if precision == "bf16" else "f32", frame0, 0)
            idx = np.arange(len(snr_db), dtype=np.int32)
            out = np.zeros((len(snr_db), 4), dtype=np.uint64)
            check(self.lib.nvd_mc_sweep(self.ctx, C.byref(cfg), snr_db.ctypes.data_as(C.POINTER(C.c_float)),
                                        idx.ctypes.data_as(C.POINTER(C.c_int32)), len(snr_db), int(min_frame_errors),
                                        int(max_frames), int(chunk_frames), out.ctypes.data))
            return out
        total = np.zeros((len(snr_db), 4), dtype=np.uint64)
        for i, db in enumerate(snr_db):
            done = 0
            while done < max_frames and total[i, 1] < min_frame_errors:
                n = int(min(chunk_frames, max_frames - done))
                c = self.run(h, code, L, [db], n, frame0 + done, mode, precision, None, [i], rank, world, all_reduce)
                total[i] += c[0]
                done += n
        return total


def _code(code) -> int:
    if isinstance(code, str):
        return _lib.CODES[code]
    return int(code)


def ber_fer(counters: np.ndarray):
    c = np.asarray(counters, dtype=np.float64)
    return c[..., 0] / np.maximum(c[..., 3], 1), c[..., 1] / np.maximum(c[..., 2], 1)


def wilson(errors: float, trials: float, z: float = 1.96):
    if trials <= 0:
        return 0.0, 1.0
    p = errors / trials
    d = 1 + z * z / trials
    c = p + z * z / (2 * trials)
    hw = z * math.sqrt(p * (1 - p) / trials + z * z / (4 * trials * trials))
    return max(0.0, (c - hw) / d), min(1.0, (c + hw) / d)
