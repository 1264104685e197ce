"""Philox4x32-10 and the RANSAC hypothesis stream -- TEST INFRASTRUCTURE (oracle side).

This file is part of ``oracle/``: it may be imported only by ``tests/``,
``__graft_entry__.smoke()`` and the CPU-baseline legs of ``bench.py``.  The
product (``mvoscalerecovery_b200``) carries its own device implementation in
``csrc/philox.cuh``; the two are written independently and pinned against the
published Random123 known-answer vectors.

Why a stream has to be *defined* at all: the reference draws its RANSAC samples
with CPython ``random.sample`` after ``random.seed(None)`` (OS entropy) on every
call (``src/thirdparty/Ransac/ransac.py:6,10``), so it is not reproducible.
Parity is only definable by injecting a sampler; this is that sampler.

Stream definition (shared by oracle, reference harness and CUDA kernel)
----------------------------------------------------------------------
  key     = (seed & 0xffffffff, seed >> 32)
  counter = (hypothesis index, frame index in sequence, sequence id, 0)
  (r0, r1, r2, _) = Philox4x32-10(counter, key)
  N = len(data)                      # number of *positions* in the vertex list
  i0 = (r0 * N)     >> 32
  i1 = (r1 * (N-1)) >> 32 ; if i1 >= i0: i1 += 1
  i2 = (r2 * (N-2)) >> 32 ; for a in sorted((i0, i1)): if i2 >= a: i2 += 1
  sample = [data[i0], data[i1], data[i2]]      # three distinct positions
"""
import numpy as np

M0 = 0xD2511F53
M1 = 0xCD9E8D57
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32_10(ctr, key):
    """Scalar Philox4x32-10 on Python ints. ctr: 4 words, key: 2 words -> 4 words."""
    c0, c1, c2, c3 = [int(c) & MASK for c in ctr]
    k0, k1 = [int(k) & MASK for k in key]
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> 32, p0 & MASK
        hi1, lo1 = p1 >> 32, p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & MASK, lo1, (hi0 ^ c3 ^ k1) & MASK, lo0
        k0 = (k0 + W0) & MASK
        k1 = (k1 + W1) & MASK
    return c0, c1, c2, c3


def philox4x32_10_np(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10 on uint64-held 32-bit words (numpy arrays or scalars)."""
    c0 = np.asarray(c0, dtype=np.uint64)
    c1 = np.asarray(c1, dtype=np.uint64)
    c2 = np.asarray(c2, dtype=np.uint64)
    c3 = np.asarray(c3, dtype=np.uint64)
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint64(k0)
    k1 = np.uint64(k1)
    m = np.uint64(MASK)
    s32 = np.uint64(32)
    for _ in range(10):
        p0 = np.uint64(M0) * c0
        p1 = np.uint64(M1) * c2
        hi0, lo0 = p0 >> s32, p0 & m
        hi1, lo1 = p1 >> s32, p1 & m
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & m, lo1, (hi0 ^ c3 ^ k1) & m, lo0
        k0 = (k0 + np.uint64(W0)) & m
        k1 = (k1 + np.uint64(W1)) & m
    return c0, c1, c2, c3


def sample3_positions(seed, hyp, frame, seq, n):
    """Three distinct positions in range(n) for hypothesis ``hyp`` of ``frame`` (scalar)."""
    r0, r1, r2, _ = philox4x32_10((hyp, frame, seq, 0), (seed & MASK, (seed >> 32) & MASK))
    i0 = (r0 * n) >> 32
    i1 = (r1 * (n - 1)) >> 32
    if i1 >= i0:
        i1 += 1
    i2 = (r2 * (n - 2)) >> 32
    for a in sorted((i0, i1)):
        if i2 >= a:
            i2 += 1
    return i0, i1, i2


def sample3_positions_np(seed, hyps, frame, seq, n):
    """Vectorised over hypothesis indices. Returns (H,3) int64."""
    hyps = np.asarray(hyps, dtype=np.uint64)
    r0, r1, r2, _ = philox4x32_10_np(hyps, np.uint64(frame), np.uint64(seq), np.uint64(0),
                                     seed & MASK, (seed >> 32) & MASK)
    n = np.uint64(n)
    s32 = np.uint64(32)
    i0 = ((r0 * n) >> s32).astype(np.int64)
    i1 = ((r1 * (n - np.uint64(1))) >> s32).astype(np.int64)
    i1 = i1 + (i1 >= i0)
    i2 = ((r2 * (n - np.uint64(2))) >> s32).astype(np.int64)
    lo = np.minimum(i0, i1)
    hi = np.maximum(i0, i1)
    i2 = i2 + (i2 >= lo)
    i2 = i2 + (i2 >= hi)
    return np.stack([i0, i1, i2], axis=1)


class PhiloxShim:
    """Drop-in for the ``random`` module global of the reference's ransac.py.

    ``run_ransac`` resolves ``random`` in its module globals
    (src/thirdparty/Ransac/ransac.py:1,6,10), so the harness assigns an instance
    of this class there.  ``seed`` is a no-op; every ``sample`` consumes the next
    hypothesis index of the frame set through ``begin_frame``.
    """

    def __init__(self, seed):
        self.seed_value = int(seed)
        self.frame = 0
        self.seq = 0
        self.hyp = 0
        self.log = []          # (hyp, i0, i1, i2) per sample call of the current frame

    def begin_frame(self, frame, seq=0):
        self.frame = int(frame)
        self.seq = int(seq)
        self.hyp = 0
        self.log = []

    def seed(self, _x=None):
        return None

    def sample(self, population, k):
        assert int(k) == 3, "the plane model draws 3 positions"
        n = len(population)
        i0, i1, i2 = sample3_positions(self.seed_value, self.hyp, self.frame, self.seq, n)
        self.log.append((self.hyp, i0, i1, i2))
        self.hyp += 1
        return [population[i0], population[i1], population[i2]]
