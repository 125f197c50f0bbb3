"""Oracle: Philox4x32-10 counter-based RNG (TEST INFRASTRUCTURE).

Not a reference function: the reference draws noise from numpy's unseeded global
Mersenne Twister (`np.random.normal`, /root/reference/src/fnn_viterbi.py:250), which is
irreproducible.  The B200 path replaces it by Philox4x32-10 (Salmon et al., SC'11,
"Parallel random numbers: as easy as 1, 2, 3"; Random123 v1.14 constants) so that
every frame is a pure function of (seed, SNR index, frame index).  This module is the
independent numpy statement the CUDA generator (csrc/philox.cuh) is checked against,
bit-exactly, including the Random123 known-answer vectors (tests/test_philox.py).

Counter / key convention shared with the CUDA side (csrc/frame_gen.cuh):
    key     = (seed & 0xffffffff, seed >> 32)
    counter = (frame & 0xffffffff, frame >> 32, block, (snr_idx << 8) | stream)
    stream  = STREAM_NOISE (1) for channel noise, STREAM_INFO (2) for source bits
One call yields 4 words = the noise of symbols 4*block .. 4*block+3 (or 128 source bits).
"""
from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)
STREAM_NOISE = 1
STREAM_INFO = 2
DEFAULT_SEED = 20170710


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10. All inputs broadcastable uint32 arrays; returns 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = p0.astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = p1.astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32(k0 + W0)
            k1 = np.uint32(k1 + W1)
    return c0, c1, c2, c3


def frame_words(seed: int, stream_word: int, frames: np.ndarray, block: int):
    """Philox words for (frame, block) under the shared counter convention."""
    f = np.asarray(frames, dtype=np.uint64)
    return philox4x32_10(
        (f & np.uint64(0xFFFFFFFF)).astype(np.uint32),
        (f >> np.uint64(32)).astype(np.uint32),
        np.uint32(block),
        np.uint32(stream_word),
        seed & 0xFFFFFFFF,
        (seed >> 32) & 0xFFFFFFFF,
    )


def box_muller(a: np.ndarray, b: np.ndarray):
    """Two uint32 words -> two N(0,1) draws, float64 statement of csrc/frame_gen.cuh:
        u1 = (a + 0.5) * 2^-32   in (0,1)
        th = int32(b) * pi * 2^-31   in [-pi, pi)
        r  = sqrt(-2 ln u1);  z0 = r cos th;  z1 = r sin th
    """
    u1 = (a.astype(np.float64) + 0.5) * 2.0 ** -32
    th = b.astype(np.int32).astype(np.float64) * (np.pi * 2.0 ** -31)
    r = np.sqrt(-2.0 * np.log(u1))
    return r * np.cos(th), r * np.sin(th)


def noise_normals(seed: int, snr_idx: int, frames: np.ndarray, n: int) -> np.ndarray:
    """N(0,1) noise matrix [len(frames), n] (float64) for the given global frame indices."""
    frames = np.asarray(frames, dtype=np.uint64)
    out = np.empty((frames.shape[0], 4 * ((n + 3) // 4)), dtype=np.float64)
    sw = ((snr_idx & 0xFFFFFF) << 8) | STREAM_NOISE
    for blk in range((n + 3) // 4):
        x0, x1, x2, x3 = frame_words(seed, sw, frames, blk)
        z0, z1 = box_muller(x0, x1)
        z2, z3 = box_muller(x2, x3)
        out[:, 4 * blk + 0] = z0
        out[:, 4 * blk + 1] = z1
        out[:, 4 * blk + 2] = z2
        out[:, 4 * blk + 3] = z3
    return out[:, :n]


def info_bits(seed: int, frames: np.ndarray, L: int) -> np.ndarray:
    """Bernoulli(1/2) source bits [len(frames), L] from stream STREAM_INFO (bit i of the
    little-endian word stream x0,x1,x2,x3 of block i//128)."""
    frames = np.asarray(frames, dtype=np.uint64)
    nblk = (L + 127) // 128
    words = np.empty((frames.shape[0], 4 * nblk), dtype=np.uint32)
    for blk in range(nblk):
        x = frame_words(seed, STREAM_INFO, frames, blk)
        for j in range(4):
            words[:, 4 * blk + j] = x[j]
    by = words.view(np.uint8)  # little-endian host
    bits = np.unpackbits(by, axis=1, bitorder="little")
    return bits[:, :L].astype(np.uint8)
