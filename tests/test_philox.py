"""Known-answer tests for the Philox4x32-10 stream shared by oracle/philox.py and the device RNG."""
import numpy as np

from oracle import philox


def _one(c, k):
    out = philox.philox4x32_10(*[np.array([x], dtype=np.uint32) for x in c], k[0], k[1])
    return [int(o[0]) for o in out]


def test_random123_known_answers():
    # Random123 kat_vectors: philox4x32 10 rounds
    assert _one((0, 0, 0, 0), (0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert _one((0xffffffff,) * 4, (0xffffffff, 0xffffffff)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert _one((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_normal_stream_properties():
    z = philox.normal_noise(seed=42, it=3, n_samples=4001, n_pairs=6)
    assert z.shape == (4001, 6) and z.dtype == np.float64
    assert abs(z.mean()) < 0.03 and abs(z.std() - 1.0) < 0.03
    z2 = philox.normal_noise(seed=42, it=3, n_samples=4001, n_pairs=6)
    assert np.array_equal(z, z2)
    assert not np.array_equal(z, philox.normal_noise(seed=42, it=4, n_samples=4001, n_pairs=6))
    # column j depends only on (seed, it, j): sharding pairs across ranks reproduces the same noise
    assert np.array_equal(z[:, 2:4], philox.normal_noise(42, 3, 4001, 6)[:, 2:4])
