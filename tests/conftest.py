"""pytest configuration: path set-up and the `gpu` marker.

`-m "not gpu"` runs on the CPU-only build container (oracle vs golden vectors, host logic,
C-ABI symbol check, gloo world_size-2 sharding).  `-m gpu` runs on a B200 and calls through
the C-ABI (libnvd.so via ctypes).  Nothing here reads /root/reference at run time.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "neural-network-based-viterbi-decoding_b200")
for p in (ROOT, PKG, os.path.join(PKG, "src")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def lib():
    """The ctypes binding of libnvd.so (fails loudly when the library is not built)."""
    from nvd import _lib
    return _lib


@pytest.fixture(scope="session")
def golden_fwd():
    """Outputs of the reference's own decoder graph functions run under the numpy `tensorflow`
    stand-in (tools/gen_golden_forward.py)."""
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_forward.npz"), allow_pickle=False)
