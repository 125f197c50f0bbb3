"""Philox4x32-10 known-answer tests (Random123 kat_vectors) for the oracle generator."""
import numpy as np

from oracle import philox


def _run(ctr, key):
    return [int(v) for v in philox.philox4x32_10(*[np.uint32(c) for c in ctr], key[0], key[1])]


def test_random123_kats():
    assert _run([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert _run([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert _run([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == \
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_normals_are_standard_and_reproducible():
    f = np.arange(1 << 16, dtype=np.uint64)
    z = philox.noise_normals(philox.DEFAULT_SEED, 3, f, 14)
    assert z.shape == (1 << 16, 14)
    assert abs(z.mean()) < 5e-3 and abs(z.std() - 1.0) < 5e-3
    assert abs((z ** 4).mean() - 3.0) < 0.05
    z2 = philox.noise_normals(philox.DEFAULT_SEED, 3, f[1000:1010], 14)
    assert np.array_equal(z[1000:1010], z2)                       # pure function of the frame index
    z3 = philox.noise_normals(philox.DEFAULT_SEED, 4, f[:10], 14)
    assert not np.array_equal(z[:10], z3)                         # SNR index decorrelates


def test_info_bits_balanced():
    b = philox.info_bits(philox.DEFAULT_SEED, np.arange(4096, dtype=np.uint64), 518)
    assert b.shape == (4096, 518) and abs(b.mean() - 0.5) < 5e-3


def test_six_normals_per_block_are_gaussian_and_independent():
    """One Philox call feeds three Box-Muller pairs (27-bit radius, 15-bit angle each): marginals, tails,
    independence within a pair / across pairs of a block, and across blocks."""
    from scipy import stats
    f = np.arange(1 << 18, dtype=np.uint64)
    z = philox.noise_normals(philox.DEFAULT_SEED, 1, f, 18)               # three whole blocks
    assert stats.kstest(z[:, 0], "norm").pvalue > 1e-3 and stats.kstest(z[:, 5], "norm").pvalue > 1e-3
    n = z.size
    for t, p in ((2.0, 4.55e-2), (3.0, 2.70e-3), (4.0, 6.33e-5)):
        k = (np.abs(z) > t).sum()
        assert abs(k - n * p) < 5 * np.sqrt(n * p)                        # two-sided tail mass
    c = np.corrcoef(z.T)
    assert np.abs(c - np.eye(18)).max() < 1e-2                            # symbols of a frame are uncorrelated
    assert abs(np.corrcoef(z[:, 0] ** 2, z[:, 1] ** 2)[0, 1]) < 1e-2      # ... and the pair's squares (radius shared!)
    # the three pairs of a block share word 3 only through disjoint bit fields
    x = philox.frame_words(philox.DEFAULT_SEED, (1 << 8) | philox.STREAM_NOISE, f[:8], 0)
    zz = philox.block_normals(*x)
    assert np.allclose(zz, z[:8, :6])
    assert np.abs(z).max() < 6.24                                         # 27-bit radius: sqrt(2 * 28 * ln 2)
