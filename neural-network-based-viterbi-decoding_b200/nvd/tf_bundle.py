"""TensorFlow checkpoint format V2 ("tensor bundle") without TensorFlow: the files `tf.train.Saver.save(sess, prefix)`
writes and `saver.restore(sess, prefix)` reads in the reference (/root/reference/src/fnn_viterbi_mixed_snr.py:742,
782-784,808-811; fnn_viterbi_bitwise.py:470,500; fnn_viterbi_seq_predict.py:494-503,587) — SURVEY.md §8 row f4.

    <prefix>.index                  an SSTable (the LevelDB table format of tensorflow/core/lib/io/table*.cc):
                                    key "" -> BundleHeaderProto, key <variable name> -> BundleEntryProto
    <prefix>.data-00000-of-00001    the tensors' bytes, little-endian, back to back at the entries' offsets

What is stated here follows the published format (tensorflow/core/util/tensor_bundle/tensor_bundle.cc,
core/protobuf/tensor_bundle.proto, core/lib/io/format.cc, block_builder.cc): blocks of prefix-compressed entries with a
restart array, a 5-byte trailer (compression type + masked CRC-32C) behind every block, a 48-byte footer with the
meta-index and index handles and the magic number 0xdb4775248b80fb57.  TensorFlow is not installable in this
environment and the reference ships no checkpoint, so the reader has been exercised only against files from the
writer below (tests/test_tf_bundle.py) — it verifies every checksum it meets and refuses what it does not know
(compressed blocks, sliced or string tensors) instead of guessing.
"""
from __future__ import annotations

import os
import struct

import numpy as np

MAGIC = 0xDB4775248B80FB57
FOOTER_BYTES = 48
RESTART_INTERVAL = 16

# tensorflow/core/framework/types.proto
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
           17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
_DT_OF = {np.dtype(v): k for k, v in _DTYPES.items()}


# ---- CRC-32C (Castagnoli), slicing by 8, and LevelDB's mask -------------------------------------------------------
def _crc_tables():
    t0 = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        t0.append(c)
    tables = [t0]
    for k in range(1, 8):
        prev = tables[-1]
        tables.append([(prev[i] >> 8) ^ t0[prev[i] & 0xFF] for i in range(256)])
    return tables


_T = _crc_tables()


def crc32c(data: bytes, crc: int = 0) -> int:
    crc ^= 0xFFFFFFFF
    mv = memoryview(data)
    n8 = len(mv) // 8 * 8
    t0, t1, t2, t3, t4, t5, t6, t7 = _T
    for lo, hi in struct.iter_unpack("<II", mv[:n8]):
        lo ^= crc
        crc = (t7[lo & 0xFF] ^ t6[(lo >> 8) & 0xFF] ^ t5[(lo >> 16) & 0xFF] ^ t4[lo >> 24] ^
               t3[hi & 0xFF] ^ t2[(hi >> 8) & 0xFF] ^ t1[(hi >> 16) & 0xFF] ^ t0[hi >> 24])
    for b in mv[n8:]:
        crc = (crc >> 8) ^ t0[(crc ^ b) & 0xFF]
    return crc ^ 0xFFFFFFFF


def mask_crc(crc: int) -> int:
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


# ---- varints and the two protos -------------------------------------------------------------------------------------
def _put_varint(v: int) -> bytes:
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _get_varint(buf, pos: int):
    v = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        v |= (b & 0x7F) << shift
        if not b & 0x80:
            return v, pos
        shift += 7
        if shift > 63:
            raise ValueError("varint too long")


def _fields(buf):
    """(field number, wire type, value) of a serialized message; length-delimited values as bytes."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _get_varint(buf, pos)
        f, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v, pos = struct.unpack_from("<Q", buf, pos)[0], pos + 8
        elif wt == 2:
            ln, pos = _get_varint(buf, pos)
            v, pos = bytes(buf[pos:pos + ln]), pos + ln
        elif wt == 5:
            v, pos = struct.unpack_from("<I", buf, pos)[0], pos + 4
        else:
            raise ValueError(f"unsupported wire type {wt}")
        yield f, wt, v


def _encode_header() -> bytes:
    # BundleHeaderProto { num_shards = 1; endianness = LITTLE (0, default: omitted); version { producer = 1 } }
    return b"\x08\x01" + b"\x1a\x02\x08\x01"


def _encode_entry(dtype: int, shape, offset: int, size: int, crc: int) -> bytes:
    # BundleEntryProto { dtype = 1; shape = 2 { dim = 2 { size = 1 } }; shard_id = 3; offset = 4; size = 5; crc32c = 6 (fixed32) }
    dims = b"".join(b"\x12" + _put_varint(len(d)) + d for d in (b"\x08" + _put_varint(int(s)) for s in shape))
    out = b"\x08" + _put_varint(dtype) + b"\x12" + _put_varint(len(dims)) + dims
    if offset:
        out += b"\x20" + _put_varint(offset)
    out += b"\x28" + _put_varint(size) + b"\x35" + struct.pack("<I", crc)
    return out


def _decode_entry(buf):
    e = {"dtype": 0, "shape": [], "shard_id": 0, "offset": 0, "size": 0, "crc32c": None, "slices": 0}
    for f, _, v in _fields(buf):
        if f == 1:
            e["dtype"] = v
        elif f == 2:
            for f2, _, v2 in _fields(v):
                if f2 == 2:
                    size = 0
                    for f3, _, v3 in _fields(v2):
                        if f3 == 1:
                            size = v3 - (1 << 64) if v3 >> 63 else v3
                    e["shape"].append(size)
                elif f2 == 3 and v2:
                    raise ValueError("tensor of unknown rank")
        elif f == 3:
            e["shard_id"] = v
        elif f == 4:
            e["offset"] = v
        elif f == 5:
            e["size"] = v
        elif f == 6:
            e["crc32c"] = v
        elif f == 7:
            e["slices"] += 1
    return e


# ---- SSTable blocks ----------------------------------------------------------------------------------------------------
def _build_block(items, restart_interval: int) -> bytes:
    out, restarts, last, count = bytearray(), [], b"", 0
    for key, value in items:
        if count % restart_interval == 0:
            restarts.append(len(out))
            shared = 0
        else:
            shared = 0
            while shared < min(len(last), len(key)) and last[shared] == key[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(value)) + key[shared:] + value
        last = key
        count += 1
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def _read_block(buf, offset: int, size: int):
    body = bytes(buf[offset:offset + size])
    ctype = buf[offset + size]
    stored = struct.unpack_from("<I", buf, offset + size + 1)[0]
    if mask_crc(crc32c(body + bytes([ctype]))) != stored:
        raise ValueError("index block checksum mismatch")
    if ctype != 0:
        raise ValueError("compressed index blocks are not supported (compression type %d)" % ctype)
    n_restarts = struct.unpack_from("<I", body, len(body) - 4)[0]
    end = len(body) - 4 - 4 * n_restarts
    pos, key, items = 0, b"", []
    while pos < end:
        shared, pos = _get_varint(body, pos)
        non_shared, pos = _get_varint(body, pos)
        vlen, pos = _get_varint(body, pos)
        key = key[:shared] + body[pos:pos + non_shared]
        pos += non_shared
        items.append((key, body[pos:pos + vlen]))
        pos += vlen
    return items


def _handle(offset: int, size: int) -> bytes:
    return _put_varint(offset) + _put_varint(size)


# ---- public API ----------------------------------------------------------------------------------------------------------
def write_bundle(prefix: str, tensors: dict) -> None:
    """Write {variable name: array} as <prefix>.index and <prefix>.data-00000-of-00001 (one shard, no compression)."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    entries, offset = [(b"", _encode_header())], 0
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        for name in sorted(tensors, key=lambda s: s.encode()):
            a = np.asarray(tensors[name])
            if not a.flags.c_contiguous:
                a = np.ascontiguousarray(a)            # (not for 0-d arrays: ascontiguousarray would make them 1-d)
            if a.dtype.byteorder == ">":
                a = a.astype(a.dtype.newbyteorder("<"))
            if np.dtype(a.dtype) not in _DT_OF:
                raise TypeError(f"{name}: dtype {a.dtype} has no TensorFlow counterpart here")
            raw = a.tobytes()
            f.write(raw)
            entries.append((name.encode(), _encode_entry(_DT_OF[np.dtype(a.dtype)], a.shape, offset, len(raw), mask_crc(crc32c(raw)))))
            offset += len(raw)
    if any(k == b"" for k, _ in entries[1:]):
        raise ValueError("empty variable name")
    out = bytearray()

    def emit(block: bytes):
        off = len(out)
        out.extend(block)
        out.extend(b"\x00" + struct.pack("<I", mask_crc(crc32c(block + b"\x00"))))
        return off, len(block)

    data = emit(_build_block(entries, RESTART_INTERVAL))            # a few dozen variables: one data block
    meta = emit(_build_block([], RESTART_INTERVAL))
    index = emit(_build_block([(entries[-1][0] + b"\x00", _handle(*data))], 1))      # a key >= the block's last key
    footer = _handle(*meta) + _handle(*index)
    out.extend(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", MAGIC))
    with open(prefix + ".index", "wb") as f:
        f.write(bytes(out))


def read_bundle(prefix: str, verify: bool = True) -> dict:
    """{variable name: array} of the checkpoint `prefix` (as passed to Saver.save / Saver.restore)."""
    with open(prefix + ".index", "rb") as f:
        buf = f.read()
    if len(buf) < FOOTER_BYTES or struct.unpack_from("<Q", buf, len(buf) - 8)[0] != MAGIC:
        raise ValueError(f"{prefix}.index is not a TensorFlow tensor-bundle index (bad magic number)")
    foot = buf[len(buf) - FOOTER_BYTES:]
    pos = 0
    _, pos = _get_varint(foot, pos)              # meta-index handle (unused by bundles)
    _, pos = _get_varint(foot, pos)
    ioff, pos = _get_varint(foot, pos)
    isize, pos = _get_varint(foot, pos)
    entries = {}
    num_shards = 1
    for _, handle in _read_block(buf, ioff, isize):
        doff, p2 = _get_varint(handle, 0)
        dsize, _ = _get_varint(handle, p2)
        for key, value in _read_block(buf, doff, dsize):
            if key == b"":
                for f_, _, v in _fields(value):
                    if f_ == 1:
                        num_shards = v
                    elif f_ == 2 and v != 0:
                        raise ValueError("big-endian bundle")
            else:
                entries[key.decode()] = _decode_entry(value)
    shards = {}
    out = {}
    for name, e in entries.items():
        if e["slices"]:
            raise ValueError(f"{name}: partitioned (sliced) variables are not supported")
        if e["dtype"] not in _DTYPES:
            raise ValueError(f"{name}: TensorFlow dtype {e['dtype']} is not supported")
        sid = e["shard_id"]
        if sid not in shards:
            with open(f"{prefix}.data-{sid:05d}-of-{num_shards:05d}", "rb") as f:
                shards[sid] = f.read()
        raw = shards[sid][e["offset"]:e["offset"] + e["size"]]
        dt = np.dtype(_DTYPES[e["dtype"]])
        if len(raw) != e["size"] or e["size"] != int(np.prod(e["shape"], dtype=np.int64)) * dt.itemsize:
            raise ValueError(f"{name}: size {e['size']} does not match shape {e['shape']} of {dt}")
        if verify and e["crc32c"] is not None and mask_crc(crc32c(raw)) != e["crc32c"]:
            raise ValueError(f"{name}: tensor checksum mismatch")
        out[name] = np.frombuffer(raw, dtype=dt.newbyteorder("<")).reshape(e["shape"]).astype(dt)
    return out


def is_bundle(prefix: str) -> bool:
    return os.path.exists(prefix + ".index")
