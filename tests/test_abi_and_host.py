"""CPU-only checks: the C-ABI library loads and exports every symbol include/nvd.h declares (no
compute call), the product path fails loudly without a GPU, and the host-side helpers."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import encoders as oenc, channel as ochan

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    if not os.path.exists(g.LIB):
        g.build()
    return g.LIB


def test_library_exports_every_header_symbol(built):
    from nvd import _lib
    lib = ctypes.CDLL(built)
    names = _lib.header_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), n
        assert n in _lib._PROTOS, f"ctypes prototype missing for {n}"
    assert set(_lib._PROTOS) == set(names)
    loaded = _lib.load()
    assert loaded.nvd_version() == 100


def test_library_contains_blackwell_tensor_core_code(built):
    """SASS evidence that the decoder is a tcgen05 kernel (UTCHMMA / LDTM / UBLKCP), sm_100a only."""
    r = subprocess.run(["cuobjdump", "-sass", built], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in r.stdout
    for mnemonic in ("UTCHMMA", "LDTM", "STTM", "UBLKCP", "MUFU.TANH", "FFMA2"):
        assert mnemonic in r.stdout, mnemonic
    assert "HMMA.16" not in r.stdout            # no legacy mma.sync path


def test_no_gpu_means_loud_failure(built):
    from nvd import _lib, montecarlo, NvdError
    if _lib.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(NvdError):
        montecarlo.Simulator(device=0)


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "neural-network-based-viterbi-decoding_b200")
    for base, _, files in os.walk(pkg):
        if "build" in base:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, f), errors="ignore").read()
                assert "import oracle" not in text and "from oracle" not in text, os.path.join(base, f)


def test_host_helpers():
    from nvd import montecarlo as mc
    x = np.random.default_rng(0).standard_normal(1000).astype(np.float32)
    assert np.array_equal(mc.bf16_to_f32(mc.f32_to_bf16(x)), ochan.bf16_round(x))
    assert np.array_equal(mc.class_to_packed(np.arange(128), 7), oenc.class_to_packed(np.arange(128), 7))
    assert np.allclose(mc.snr_db_to_sigma(np.arange(-2, 10)), ochan.snr_db_to_sigma(np.arange(-2, 10)), rtol=1e-15)
    for frames, world in [(100, 3), (1 << 24, 8), (7, 8), (0, 2)]:
        parts = [mc.shard_range(5, frames, r, world) for r in range(world)]
        assert parts[0][0] == 5 and sum(n for _, n in parts) == frames
        for (lo, n), (lo2, _) in zip(parts[:-1], parts[1:]):
            assert lo + n == lo2
    lo, hi = mc.wilson(10, 1000)
    assert lo < 0.01 < hi


def test_weights_roundtrip(tmp_path):
    from nvd import weights
    m = weights.init_mlp([14, 128, 64, 32, 128], ["tanh"] * 4, seed=1)
    assert [W.shape for W in m.Ws] == [(14, 128), (128, 64), (64, 32), (32, 128)]
    assert all(np.abs(W).max() <= 2 * np.sqrt(2.0 / W.shape[0]) + 1e-6 for W in m.Ws)   # truncated at 2 sigma
    w, b = weights.to_dicts(m)
    assert set(w) == {"h1", "h2", "h3", "out"} and set(b) == {"b1", "b2", "b3", "out"}
    p = str(tmp_path / "m.npz")
    weights.save_mlp(p, m)
    z = np.load(p)
    assert {"weights_h1", "weights_out", "bias_h1", "bias_out"} <= set(z.files)          # TF variable names
    m2 = weights.load_mlp(p)
    assert all(np.array_equal(a, c) for a, c in zip(m.Ws, m2.Ws)) and m2.acts == [1, 1, 1, 1]
    m3 = weights.from_dicts(w, b, ["tanh"] * 4)
    assert all(np.array_equal(a, c) for a, c in zip(m.Ws, m3.Ws))
