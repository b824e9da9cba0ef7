"""Counter-based random streams of the native sampler (csrc/sampler.cuh), restated in numpy.

The native batched sampler (`BatchedAPMSampler(rng='native')`) cannot reproduce numpy's Mersenne-Twister streams on the
device, so it defines its own: Philox4x32-10 (Salmon et al., SC'11; the constants and round function below are the
published ones, checked against the Random123 known-answer vectors in tests/test_philox.py) keyed by the chain's seed.
This module is the host mirror of those streams: `PhiloxStream` offers the three calls the chain generators of
`apm_b200.batched` make on a `numpy.random.RandomState` -- `uniform()`, `normal()` / `normal(size=k)` and
`normal(size=(n, N))` -- so that `BatchedAPMSampler(rng='philox')` runs the Python (reference-pinned) chain logic on
exactly the random numbers the native sampler uses, and the two can be compared draw for draw.

Streams of one chain (key = the 64-bit seed):
  scalar draws   counter (k, 0, 0, 1), k = 0, 1, 2, ... in the order the chain asks for them:
                 uniform = u1, normal = sqrt(-2 log u1) cos(2 pi u2)
  bulk draws     the d-th (n, N) block of standard normals of the chain uses counters (i * ceil(N/2) + q, 0, d, 2):
                 element (i, 2q) = sqrt(-2 log u1) cos(2 pi u2), element (i, 2q + 1) = sqrt(-2 log u1) sin(2 pi u2)
with u1 = ((x0 | x1 << 32) >> 11 + 0.5) 2^-53 and u2 likewise from (x2, x3) of the four 32-bit outputs.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)
STREAM_SCALAR, STREAM_BULK = 1, 2


def philox4x32(counter, key, rounds=10):
    """counter: (..., 4) uint32-valued array, key: (2,) -> (..., 4) uint32 values as uint64 arrays."""
    c = [np.asarray(counter[..., j], dtype=np.uint64) for j in range(4)]
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(rounds):
        p0 = M0 * c[0]
        p1 = M1 * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c = [hi1 ^ c[1] ^ np.uint64(k0), lo1, hi0 ^ c[3] ^ np.uint64(k1), lo0]
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return np.stack(c, axis=-1)


def _u01(lo, hi):
    x = (lo | (hi << np.uint64(32))) >> np.uint64(11)
    return (x.astype(np.float64) + 0.5) * (2.0 ** -53)


def _pairs(out):
    u1 = _u01(out[..., 0], out[..., 1])
    u2 = _u01(out[..., 2], out[..., 3])
    r = np.sqrt(-2.0 * np.log(u1))
    return u1, r * np.cos(2.0 * np.pi * u2), r * np.sin(2.0 * np.pi * u2)


def seed_key(seed):
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return (seed & 0xFFFFFFFF, seed >> 32)


def bulk_normals(seed, draw, n, N):
    """The draw-th (n, N) block of standard normals of the chain with this seed."""
    half = (N + 1) // 2
    idx = np.arange(n * half, dtype=np.uint64)
    ctr = np.zeros((n * half, 4), dtype=np.uint64)
    ctr[:, 0] = idx & MASK
    ctr[:, 1] = idx >> np.uint64(32)
    ctr[:, 2] = draw
    ctr[:, 3] = STREAM_BULK
    _, z0, z1 = _pairs(philox4x32(ctr, seed_key(seed)))
    z = np.stack([z0, z1], axis=-1).reshape(n, 2 * half)
    return np.ascontiguousarray(z[:, :N])


class PhiloxStream(object):
    """The subset of numpy.random.RandomState the chain generators use, on the native sampler's streams."""

    def __init__(self, seed):
        self.seed = int(seed)
        self.key = seed_key(seed)
        self.n_scalar = 0
        self.n_bulk = 0

    def _scalar(self):
        ctr = np.array([[self.n_scalar & 0xFFFFFFFF, self.n_scalar >> 32, 0, STREAM_SCALAR]], dtype=np.uint64)
        self.n_scalar += 1
        u1, z0, _ = _pairs(philox4x32(ctr, self.key))
        return float(u1[0]), float(z0[0])

    def uniform(self):
        return self._scalar()[0]

    def normal(self, size=None):
        if size is None:
            return self._scalar()[1]
        if isinstance(size, tuple) and len(size) == 2:
            z = bulk_normals(self.seed, self.n_bulk, size[0], size[1])
            self.n_bulk += 1
            return z
        k = int(size[0]) if isinstance(size, tuple) else int(size)
        return np.array([self._scalar()[1] for _ in range(k)])
