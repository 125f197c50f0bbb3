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
One call yields 4 words = 128 source bits, or the noise of SIX symbols 6*block .. 6*block+5: three
Box-Muller pairs, pair i taking a 27-bit radius uniform from the top of word i and a 15-bit angle made of
the low 5 bits of word i and bits [10i, 10i+10) of word 3 (42 bits per pair; the top two bits of word 3
are unused).  27 bits put the largest radius at sqrt(2 * 28 * ln 2) = 6.23 sigma; 2^15 directions keep the
angular lattice (1.9e-4 rad) far below the bf16 rounding of the stored symbol.  (Round 1 spent a whole
32-bit word on each of radius and angle: 4 symbols per call; the Philox rounds are the dominant cost of
every Monte-Carlo kernel, so the same 128 bits now feed 6.)
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


NORMALS_PER_BLOCK = 6


def box_muller(r_bits: np.ndarray, a_bits: np.ndarray):
    """27-bit radius field, 15-bit angle field -> two N(0,1) draws; float64 statement of
    csrc/frame_gen.cuh:
        u1 = (r_bits + 0.5) * 2^-27            in (0,1)
        th = int32(a_bits << 17) * pi * 2^-31  in [-pi, pi)
        r  = sqrt(-2 ln u1);  z0 = r cos th;  z1 = r sin th
    """
    u1 = (r_bits.astype(np.float64) + 0.5) * 2.0 ** -27
    th = (a_bits.astype(np.uint32) << np.uint32(17)).astype(np.int32).astype(np.float64) * (np.pi * 2.0 ** -31)
    r = np.sqrt(-2.0 * np.log(u1))
    return r * np.cos(th), r * np.sin(th)


def block_normals(x0, x1, x2, x3):
    """Philox words of one call -> [.., 6] normals (symbols 6*block .. 6*block+5)."""
    out = []
    for i, xi in enumerate((x0, x1, x2)):
        r_bits = xi >> np.uint32(5)
        a_bits = ((xi & np.uint32(31)) << np.uint32(10)) | ((x3 >> np.uint32(10 * i)) & np.uint32(1023))
        out.extend(box_muller(r_bits, a_bits))
    return np.stack(out, axis=-1)


def noise_normals(seed: int, snr_idx: int, frames: np.ndarray, n: int) -> np.ndarray:
    """N(0,1) noise matrix [len(frames), n] (float64) for the given global frame indices."""
    frames = np.asarray(frames, dtype=np.uint64)
    nblk = (n + NORMALS_PER_BLOCK - 1) // NORMALS_PER_BLOCK
    out = np.empty((frames.shape[0], NORMALS_PER_BLOCK * nblk), dtype=np.float64)
    sw = ((snr_idx & 0xFFFFFF) << 8) | STREAM_NOISE
    for blk in range(nblk):
        out[:, 6 * blk:6 * blk + 6] = block_normals(*frame_words(seed, sw, frames, blk))
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
