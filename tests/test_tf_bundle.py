"""TensorFlow checkpoint V2 ("tensor bundle") reader / writer (nvd/tf_bundle.py): checksums against the
published CRC-32C answers, the table format field by field, a round trip of a decoder's variables (with optimizer slots
beside them, as a Saver writes), load_mlp falling back to a checkpoint, and loud failures on damaged files."""
import os
import struct
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "neural-network-based-viterbi-decoding_b200"))
from nvd import tf_bundle as tb, weights  # noqa: E402


def test_crc32c_known_answers():
    # RFC 3720 B.4 and the LevelDB / TensorFlow crc32c_test vectors
    assert tb.crc32c(b"123456789") == 0xE3069283
    assert tb.crc32c(bytes(32)) == 0x8A9136AA
    assert tb.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43
    assert tb.crc32c(bytes(range(32))) == 0x46DD794E
    assert tb.crc32c(bytes(range(31, -1, -1))) == 0x113FDB5C
    assert tb.crc32c(b"hello world") == tb.crc32c(b" world", tb.crc32c(b"hello"))          # extendable
    c = tb.crc32c(b"foo")
    assert tb.mask_crc(c) != c and tb.mask_crc(tb.mask_crc(c)) != c                           # leveldb crc32c_test "Mask"


def _model():
    return weights.init_mlp([14, 128, 64, 32, 128], ["tanh"] * 4, seed=5)


def test_round_trip_and_file_layout(tmp_path):
    m = _model()
    w, b = weights.to_dicts(m)
    tensors = {f"weights_{k}": v for k, v in w.items()}
    tensors.update({"bias_h1": b["b1"], "bias_h2": b["b2"], "bias_h3": b["b3"], "bias_out": b["out"]})
    # what a Saver adds next to the model: Adam slots, the power accumulators (scalars), an int64 step
    tensors["weights_h1/Adam"] = np.zeros_like(w["h1"])
    tensors["weights_h1/Adam_1"] = np.full_like(w["h1"], 0.25)
    tensors["beta1_power"] = np.float32(0.9)
    tensors["global_step"] = np.int64(1234567890123)
    prefix = str(tmp_path / "ckpt" / "model")
    tb.write_bundle(prefix, tensors)
    got = tb.read_bundle(prefix)
    assert sorted(got) == sorted(tensors)
    for k, v in tensors.items():
        assert got[k].dtype == np.asarray(v).dtype and got[k].shape == np.asarray(v).shape and np.array_equal(got[k], v), k

    idx = open(prefix + ".index", "rb").read()
    assert struct.unpack_from("<Q", idx, len(idx) - 8)[0] == 0xDB4775248B80FB57                 # kTableMagicNumber
    data = open(prefix + ".data-00000-of-00001", "rb").read()
    assert len(data) == sum(np.asarray(v).nbytes for v in tensors.values())                     # back to back, no padding
    # first entry of the first block: key "" -> BundleHeaderProto{num_shards: 1, version{producer: 1}}
    assert idx[:3] == bytes([0, 0, 6]) and idx[3:9] == b"\x08\x01\x1a\x02\x08\x01"
    # the tensors sit in the data file in key order
    first = sorted(tensors, key=lambda s: s.encode())[0]
    assert np.array_equal(np.frombuffer(data[:np.asarray(tensors[first]).nbytes], np.asarray(tensors[first]).dtype).reshape(np.asarray(tensors[first]).shape),
                          tensors[first])


def test_load_mlp_falls_back_to_a_checkpoint(tmp_path):
    m = _model()
    prefix = str(tmp_path / "mixed_snr_model")
    weights.save_tf_checkpoint(prefix, m)
    assert not os.path.exists(prefix + ".npz")
    for path in (prefix, prefix + ".npz"):                 # the drop-ins append .npz to the reference's path
        got = weights.load_mlp(path)
        assert got.meta["format"] == "tf-bundle"
        assert all(np.array_equal(a, b_) for a, b_ in zip(got.Ws, m.Ws)) and all(np.array_equal(a, b_) for a, b_ in zip(got.bs, m.bs))
    # an .npz beside it wins (the repo's own format carries the activations too)
    weights.save_mlp(prefix + ".npz", m)
    assert "format" not in weights.load_mlp(prefix + ".npz").meta


def test_damaged_files_fail_loudly(tmp_path):
    prefix = str(tmp_path / "m")
    tb.write_bundle(prefix, {"a": np.arange(12, dtype=np.float32).reshape(3, 4), "b": np.arange(5, dtype=np.int32)})
    idx = bytearray(open(prefix + ".index", "rb").read())
    data = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())

    bad = bytearray(data); bad[5] ^= 1
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(bad))
    with pytest.raises(ValueError, match="checksum"):
        tb.read_bundle(prefix)
    assert np.array_equal(tb.read_bundle(prefix, verify=False)["b"], np.arange(5, dtype=np.int32))
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(data[:-4]))        # truncated shard
    with pytest.raises(ValueError, match="size"):
        tb.read_bundle(prefix)
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(data))

    bad = bytearray(idx); bad[10] ^= 0x40                                       # inside the data block
    open(prefix + ".index", "wb").write(bytes(bad))
    with pytest.raises(ValueError, match="checksum"):
        tb.read_bundle(prefix)
    bad = bytearray(idx); bad[-1] ^= 1                                          # magic number
    open(prefix + ".index", "wb").write(bytes(bad))
    with pytest.raises(ValueError, match="magic"):
        tb.read_bundle(prefix)
    with pytest.raises(TypeError):
        tb.write_bundle(prefix, {"s": np.array(["text"])})


def test_many_variables_use_prefix_compression_and_restarts(tmp_path):
    # more entries than one restart interval, long shared prefixes (optimizer slots of a deep net)
    tensors = {f"layer_{i:03d}/kernel/Adam_{j}": np.full((2, 3), i + 0.5 * j, np.float32) for i in range(40) for j in range(2)}
    prefix = str(tmp_path / "many")
    tb.write_bundle(prefix, tensors)
    got = tb.read_bundle(prefix)
    assert len(got) == 80 and all(np.array_equal(got[k], v) for k, v in tensors.items())
    idx = open(prefix + ".index", "rb").read()
    assert idx.count(b"layer_") < 20                       # keys are stored as deltas: ~one full key per restart point
