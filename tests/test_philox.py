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


def test_eight_normals_per_block_are_gaussian_and_independent():
    """One Philox call feeds four Box-Muller pairs (20-bit radius, 12-bit angle each; the 16 lowest radius cells
    refined by a second call): marginals, tails, independence within a pair / across pairs of a block, across blocks."""
    from scipy import stats
    f = np.arange(1 << 18, dtype=np.uint64)
    z = philox.noise_normals(philox.DEFAULT_SEED, 1, f, 24)               # three whole blocks
    assert stats.kstest(z[:, 0], "norm").pvalue > 1e-3 and stats.kstest(z[:, 7], "norm").pvalue > 1e-3
    n = z.size
    for t, p in ((2.0, 4.55e-2), (3.0, 2.70e-3), (4.0, 6.33e-5)):
        k = (np.abs(z) > t).sum()
        assert abs(k - n * p) < 5 * np.sqrt(n * p)                        # two-sided tail mass
    c = np.corrcoef(z.T)
    assert np.abs(c - np.eye(24)).max() < 1e-2                            # symbols of a frame are uncorrelated
    assert abs(np.corrcoef(z[:, 0] ** 2, z[:, 1] ** 2)[0, 1]) < 1e-2      # ... and the pair's squares (radius shared!)
    sw = (1 << 8) | philox.STREAM_NOISE
    x = philox.frame_words(philox.DEFAULT_SEED, sw, f[:8], 0)
    t = philox.frame_words(philox.DEFAULT_SEED, (1 << 8) | philox.STREAM_NOISE_TAIL, f[:8], 0)
    zz = philox.block_normals(x, t)
    assert np.allclose(zz, z[:8, :8])


def test_far_tail_radius_is_refined():
    """Radius cells k < 16 take 32 more bits from the tail stream: the radius is continuous across the k = 15 / 16
    boundary, goes beyond the 5.4 sigma a bare 20-bit field would stop at, and the radius^2 / 2 of the tail is
    exponential like everywhere else."""
    w = np.array([0x00000FFF, 0x0000F000 | 5, 0x00010000 | 5, 0xFFFFF000], dtype=np.uint32)
    lo, _ = philox.box_muller(w, np.array([0, 0xFFFFFFFF, 0, 0], dtype=np.uint32))
    a = (w & 0xFFF).astype(np.float64) * (2 * np.pi / 4096) - np.pi
    r = lo / np.cos(a)
    assert abs(r[0] - np.sqrt(-2 * np.log(0.5 * 2.0 ** -52))) < 1e-9      # k = 0, t = 0: 8.57 sigma
    assert 0 < r[1] - r[2] < 0.01                                           # top of cell 15 meets cell 16's midpoint
    assert abs(r[3] - np.sqrt(-2 * np.log(1 - 2.0 ** -21))) < 1e-12
    # exponential tail: among pairs with k < 16 (u < 2^-16), P(u < 2^-17) = 1/2 and P(u < 2^-18) = 1/4
    rng = np.random.default_rng(5)
    k = rng.integers(0, 16, 200000).astype(np.uint32)
    ww = (k << np.uint32(12)) | rng.integers(0, 4096, 200000).astype(np.uint32)
    tt = rng.integers(0, 1 << 32, 200000, dtype=np.uint64).astype(np.uint32)
    z0, z1 = philox.box_muller(ww, tt)
    u = np.exp(-(z0 ** 2 + z1 ** 2) / 2)
    assert abs((u < 2.0 ** -17).mean() - 0.5) < 5e-3 and abs((u < 2.0 ** -18).mean() - 0.25) < 5e-3
