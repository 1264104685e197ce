"""Feature bucketing (src/detector.py:65-95).  CPU side: the oracle (oracle/extras.bucket_philox) against outputs of
the reference's own bucket() (tests/golden/bucket.npz: same cells, same order, same number of survivors per cell -- which members
survive depends on numpy's global RNG in the reference and on the Philox key here), and the SOURCE of bucket_kernel under the
pthread emulation (tests/host_sim/bucket_kernel_emu.cpp) against the oracle, index for index.  The GPU side (the kernel through
the C ABI against the oracle, index for index) is tests/test_gpu_zzz_bucket.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import extras as X

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIM = os.path.join(ROOT, "tests", "host_sim")
vp = C.c_void_p


def _p(a):
    return None if a is None else a.ctypes.data_as(vp)


def _cells(points, bs):
    return [(int(v) // bs, int(u) // bs) for u, v in np.asarray(points, dtype=np.float32).reshape(-1, 2)]


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "bucket.npz"))


def _batch(golden, sets):
    lens = [golden["f%d" % k].shape[0] for k in sets]
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    u = np.ascontiguousarray(np.concatenate([golden["f%d" % k][:, 0] for k in sets]).astype(np.float32))
    v = np.ascontiguousarray(np.concatenate([golden["f%d" % k][:, 1] for k in sets]).astype(np.float32))
    return off, u, v


def test_oracle_visits_the_reference_cells_in_the_reference_order(golden):
    for k in range(int(golden["n_sets"])):
        f, kept = golden["f%d" % k], golden["kept%d" % k]
        bs, dens = (int(x) for x in golden["par%d" % k])
        idx = X.bucket_philox(f, bs, dens, seed=3, frame=k)
        assert _cells(f[idx], bs) == _cells(kept, bs)                       # same cells, same order, same survivors per cell
        assert len(set(idx.tolist())) == idx.size                          # nothing twice
        members = {tuple(p) for p in f.tolist()}
        assert all(tuple(p) in members for p in kept.tolist())
        # a different key changes who survives, never where
        other = X.bucket_philox(f, bs, dens, seed=4, frame=k)
        assert _cells(f[other], bs) == _cells(kept, bs)
        if idx.size < f.shape[0]:
            assert not np.array_equal(idx, other)


def load_bucket_emulation():
    so, src = os.path.join(SIM, "libbucket_kernel_emu.so"), os.path.join(SIM, "bucket_kernel_emu.cpp")
    deps = [src] + [os.path.join(ROOT, "mvoscalerecovery_b200", "csrc", f) for f in ("bucket_kernel.cuh", "five_point.cuh")]
    if not os.path.isfile(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-pthread", "-Wno-unknown-pragmas", "-o", so, src])
    L = C.CDLL(so)
    L.bucket_emu.argtypes = [C.c_int32, vp, vp, vp, C.c_int32, C.c_int32, C.c_uint64, vp, C.c_int32, vp, vp, vp, C.c_int32]
    return L


def _check_against_oracle(off, u, v, bs, dens, seed, fidx, seq, index, n_out, status):
    for f in range(len(off) - 1):
        a, e = off[f], off[f + 1]
        pts = np.stack([u[a:e], v[a:e]], 1)
        bad = e - a > 4096 or not (np.isfinite(pts).all() and (pts >= 0).all())
        if bad or e == a:
            assert n_out[f] == 0 and status[f] == (0 if e == a else (2 if e - a > 4096 else 1))
            continue
        want = X.bucket_philox(pts, bs, dens, seed=seed, frame=int(fidx[f]) if fidx is not None else f, seq=seq)
        assert status[f] == 0 and n_out[f] == want.size
        assert np.array_equal(index[a:a + want.size], want), f


def test_kernel_source_on_the_host_emulation(golden):
    emu = load_bucket_emulation()
    rng = np.random.default_rng(0)
    for sets, bs, dens, grid in (((0, 1, 4), 30, 2, 2), ((2,), 50, 1, 1), ((3, 4, 0), 20, 3, 5)):
        off, u, v = _batch(golden, sets)
        fidx = np.ascontiguousarray((np.arange(len(sets)) * 7 + 2).astype(np.int32))
        index = np.full(u.size, -1, np.int32); n_out = np.full(len(sets), -1, np.int32); status = np.full(len(sets), 9, np.uint8)
        assert emu.bucket_emu(len(sets), _p(off), _p(u), _p(v), bs, dens, 2**35 + 11, _p(fidx), 4, _p(index), _p(n_out), _p(status), grid) == 0
        _check_against_oracle(off, u, v, bs, dens, 2**35 + 11, fidx, 4, index, n_out, status)
    # edge frames: empty, one feature, 4096 and 4097 features, a negative coordinate, a NaN
    lens = [0, 1, 4096, 4097, 30, 30, 5]
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    u = rng.uniform(0, 1241, off[-1]).astype(np.float32); v = rng.uniform(0, 376, off[-1]).astype(np.float32)
    u[off[4] + 3] = -1.0; v[off[5] + 7] = np.nan
    index = np.full(u.size, -1, np.int32); n_out = np.full(len(lens), -1, np.int32); status = np.full(len(lens), 9, np.uint8)
    assert emu.bucket_emu(len(lens), _p(off), _p(u), _p(v), 30, 2, 5, None, 0, _p(index), _p(n_out), _p(status), 3) == 0
    _check_against_oracle(off, u, v, 30, 2, 5, None, 0, index, n_out, status)
    assert list(status) == [0, 0, 0, 2, 1, 1, 0] and n_out[1] == 1 and n_out[2] > 900
