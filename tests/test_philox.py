"""The counter-based random streams of the native sampler (apm_b200.philox is the numpy mirror of csrc/sampler.cuh):
Philox4x32-10 against the published Random123 known-answer vectors, and the stream layout."""
import numpy as np

from apm_b200 import philox as ph


def _kat(counter, key):
    out = ph.philox4x32(np.array([counter], dtype=np.uint64), key)[0]
    return [int(x) for x in out]


def test_philox4x32_10_known_answers():
    # Random123 kat_vectors: philox4x32 10
    assert _kat([0, 0, 0, 0], (0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert _kat([0xffffffff] * 4, (0xffffffff, 0xffffffff)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert _kat([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], (0xa4093822, 0x299f31d0)) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_stream_layout_and_moments():
    s = ph.PhiloxStream(123456789012345)
    us = np.array([s.uniform() for _ in range(4000)])
    assert 0. < us.min() and us.max() < 1.
    assert abs(us.mean() - 0.5) < 0.02 and abs(us.var() - 1. / 12.) < 0.01
    zs = s.normal(size=3000)
    assert abs(zs.mean()) < 0.06 and abs(zs.std() - 1.) < 0.05
    assert s.n_scalar == 7000 and s.n_bulk == 0
    # bulk draws: one block per call, odd sample counts drop the second half of the last pair of every row
    a = s.normal(size=(50, 7))
    b = s.normal(size=(50, 7))
    assert s.n_bulk == 2 and a.shape == (50, 7) and not np.allclose(a, b)
    a8 = ph.bulk_normals(s.seed, 0, 50, 8)
    assert np.array_equal(a, a8[:, :7])
    # a chain's streams depend on its seed only
    t = ph.PhiloxStream(123456789012345)
    assert np.array_equal(np.array([t.uniform() for _ in range(10)]), us[:10])
    assert not np.array_equal(ph.bulk_normals(1, 0, 4, 4), ph.bulk_normals(2, 0, 4, 4))
    z = ph.bulk_normals(5, 3, 400, 64)
    assert abs(z.mean()) < 0.03 and abs(z.std() - 1.) < 0.03
