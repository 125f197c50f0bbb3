"""Oracle: source bits, convolutional encoders and bit packing (TEST INFRASTRUCTURE).

Restates (does not import) the reference encoders:
  encode_133171            /root/reference/src/fnn_viterbi.py:157-217
  encode_57                /root/reference/src/fnn_viterbi.py:122-154
  extract_state_bits       /root/reference/src/fnn_viterbi_bitwise.py:114-138
  encode_133171_with_state /root/reference/src/fnn_viterbi_bitwise.py:141-215
  generate_class_bits      /root/reference/src/fnn_viterbi.py:260-287

The reference hand-unrolls the first K-1 output pairs and loops over the rest; the
restatement is the equivalent zero-state, unterminated feed-forward shift register
(delay taps below), which also covers L < K-1 where the reference raises IndexError
(SURVEY.md Appendix B-5).  Output order is interleaved [g_a(i), g_b(i)].
"""
from __future__ import annotations

import numpy as np

# delay taps (i - d) of each generator, in the order the reference EMITS them
TAPS_133171 = ((0, 2, 3, 5, 6), (0, 1, 2, 3, 6))   # fnn_viterbi.py:208-215 (g=133 first, then g=171)
TAPS_57_REF = ((0, 2), (0, 1, 2))                   # fnn_viterbi.py:149-152 (g=5 first, then g=7)

CODE_57_REF = 0        # (7,5) as the reference computes it, incl. the c[1] quirk (:143)
CODE_133171 = 1
CODE_57_CAUSAL = 2     # (7,5) without the quirk

_TAPS = {CODE_57_REF: TAPS_57_REF, CODE_133171: TAPS_133171, CODE_57_CAUSAL: TAPS_57_REF}


def _as_bits(bits) -> np.ndarray:
    # reference applies `% 2` to whatever dtype arrives and stores into int32 (:181 ff.)
    a = np.asarray(bits)
    return (np.mod(a, 2)).astype(np.int64)


def conv_encode(bits, code: int = CODE_133171) -> np.ndarray:
    """Rate-1/2 zero-state unterminated encoder; bits (..., L) -> coded (..., 2L) int32."""
    b = _as_bits(bits)
    L = b.shape[-1]
    out = np.zeros(b.shape[:-1] + (2 * L,), dtype=np.int64)
    for g, taps in enumerate(_TAPS[code]):
        acc = np.zeros_like(b)
        for d in taps:
            if d < L:
                acc[..., d:] ^= b[..., : L - d]
        out[..., g::2] = acc
    if code == CODE_57_REF and L >= 2:
        # fnn_viterbi.py:143  encoded_seq[1] = (info_seq[0] + info_seq[1]) % 2  (non-causal)
        out[..., 1] ^= b[..., 1]
    return out.astype(np.int32)


def encode_133171(bits) -> np.ndarray:
    return conv_encode(bits, CODE_133171)


def encode_57(bits) -> np.ndarray:
    return conv_encode(bits, CODE_57_REF)


def extract_state_bits(state_number: int, constraint_length: int = 6) -> np.ndarray:
    """MSB-first binary of the state number (fnn_viterbi_bitwise.py:134-137).

    The reference raises for state >= 2**constraint_length; the oracle masks to the low
    bits."""
    s = int(state_number) & ((1 << constraint_length) - 1)
    return np.array([(s >> (constraint_length - 1 - i)) & 1 for i in range(constraint_length)],
                    dtype=np.int32)


def encode_133171_with_state(bits, state_num: int) -> np.ndarray:
    """State bits are prepended and the whole (L+6) sequence encoded from zero state;
    all 2(L+6) symbols are returned (fnn_viterbi_bitwise.py:162-163)."""
    b = _as_bits(bits)
    st = extract_state_bits(state_num, 6).astype(np.int64)
    return conv_encode(np.concatenate([st, b]), CODE_133171)


def generate_class_bits(block_length: int) -> np.ndarray:
    """Row i = MSB-first binary of i (np.binary_repr, fnn_viterbi.py:283)."""
    idx = np.arange(2 ** block_length, dtype=np.int64)[:, None]
    sh = np.arange(block_length - 1, -1, -1, dtype=np.int64)[None, :]
    return ((idx >> sh) & 1).astype(np.uint8)


# ----------------------------------------------------------------------------------------
# packed wire format: bit i of the frame's little-endian byte stream is the
# symbol at time i; frames are byte-aligned: ceil(L/8) info bytes, ceil(2L/8) coded bytes.
# ----------------------------------------------------------------------------------------

def pack_bits(bits: np.ndarray) -> np.ndarray:
    """(..., n) {0,1} -> (..., ceil(n/8)) uint8, time-ordered little-endian."""
    return np.packbits(np.asarray(bits).astype(np.uint8), axis=-1, bitorder="little")


def unpack_bits(packed: np.ndarray, n: int) -> np.ndarray:
    return np.unpackbits(np.asarray(packed, dtype=np.uint8), axis=-1, bitorder="little")[..., :n]


def class_to_packed(cls: np.ndarray, L: int) -> np.ndarray:
    """class index (MSB-first label of the reference) -> time-packed integer (bit t = b[t])."""
    cls = np.asarray(cls).astype(np.int64)
    out = np.zeros_like(cls)
    for t in range(L):
        out |= ((cls >> (L - 1 - t)) & 1) << t
    return out


def packed_to_class(word: np.ndarray, L: int) -> np.ndarray:
    return class_to_packed(word, L)  # bit reversal is an involution


def conv_encode_packed(info_packed: np.ndarray, L: int, code: int = CODE_133171) -> np.ndarray:
    """uint8 [B, ceil(L/8)] -> uint8 [B, ceil(2L/8)] through the bit-level oracle."""
    bits = unpack_bits(info_packed, L)
    return pack_bits(conv_encode(bits, code))
