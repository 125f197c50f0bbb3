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
One call yields 4 words = 128 source bits, or the noise of EIGHT symbols 8*block .. 8*block+7: four
Box-Muller pairs, pair p made of word p alone — a 20-bit radius field k = w >> 12 and a 12-bit angle field
a = w & 0xFFF:
    u = (k + 0.5) * 2^-20,  theta = a * 2 pi / 4096 - pi,  r = sqrt(-2 ln u),  z = (r cos theta, r sin theta).
Twenty bits alone would end the radius at 5.4 sigma on a coarse grid, so the sixteen lowest cells (k < 16,
probability 1.5e-5 per pair) are refined by 32 more bits: word p of a second call on stream STREAM_NOISE_TAIL
(same frame and block), u = (k * 2^32 + t + 0.5) * 2^-52 — radius up to 8.6 sigma, relative cell width <= 6 %
everywhere.  The CUDA side makes that second call only when a thread meets such a word.  2^12 directions keep
the angular lattice (1.5e-3 rad) below the bf16 rounding of the stored symbol.  (Round 1 spent a whole 32-bit
word on each of radius and angle: 4 symbols per call; an earlier form of round 2 used 27 + 15 bits: 6 symbols
per call.  The Philox rounds are the dominant cost of every Monte-Carlo kernel.)
"""
from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)
STREAM_NOISE = 1
STREAM_INFO = 2
STREAM_NOISE_TAIL = 3
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


NORMALS_PER_BLOCK = 8
TAIL_CELLS = 16          # radius cells k < TAIL_CELLS take 32 refinement bits


def box_muller(words: np.ndarray, tail_words: np.ndarray):
    """Philox word (20-bit radius field | 12-bit angle field) and the refinement word of the same pair -> two
    N(0,1) draws; float64 statement of csrc/frame_gen.cuh (bm_lg2_u, bm_lg2_u_tail, bm_theta)."""
    w = np.asarray(words, dtype=np.uint32)
    t = np.asarray(tail_words, dtype=np.uint32)
    k = (w >> np.uint32(12)).astype(np.float64)
    a = (w & np.uint32(0xFFF)).astype(np.float64)
    u = np.where(k < TAIL_CELLS, (k * 2.0 ** 32 + t.astype(np.float64) + 0.5) * 2.0 ** -52, (k + 0.5) * 2.0 ** -20)
    th = a * (2.0 * np.pi / 4096.0) - np.pi
    r = np.sqrt(-2.0 * np.log(u))
    return r * np.cos(th), r * np.sin(th)


def block_normals(x, t):
    """Philox words of one call (x: 4 arrays) and of the refinement call (t: 4 arrays) -> [.., 8] normals
    (symbols 8*block .. 8*block+7)."""
    out = []
    for xi, ti in zip(x, t):
        out.extend(box_muller(xi, ti))
    return np.stack(out, axis=-1)


def noise_normals(seed: int, snr_idx: int, frames: np.ndarray, n: int) -> np.ndarray:
    """N(0,1) noise matrix [len(frames), n] (float64) for the given global frame indices."""
    frames = np.asarray(frames, dtype=np.uint64)
    nblk = (n + NORMALS_PER_BLOCK - 1) // NORMALS_PER_BLOCK
    out = np.empty((frames.shape[0], NORMALS_PER_BLOCK * nblk), dtype=np.float64)
    sw = ((snr_idx & 0xFFFFFF) << 8) | STREAM_NOISE
    sw_tail = ((snr_idx & 0xFFFFFF) << 8) | STREAM_NOISE_TAIL
    for blk in range(nblk):
        out[:, 8 * blk:8 * blk + 8] = block_normals(frame_words(seed, sw, frames, blk), frame_words(seed, sw_tail, frames, blk))
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
