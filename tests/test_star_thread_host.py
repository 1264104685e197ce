"""The per-lane Delaunay star builder (csrc/gthread.cuh) without a GPU: the same __host__ __device__ source compiled by g++
(tests/host_sim/star_thread_host.cpp) over a strip-sorted set, every star it certifies compared with Qhull's -- neighbour set and
counter-clockwise order from the nearest neighbour -- on image-uniform, perspective (SURVEY 8d ground) and clustered feature sets.
Stars it cannot certify must say so (they go to the warp-per-star paths on the GPU): never a wrong ring."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
from scipy.spatial import Delaunay

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIM = os.path.join(ROOT, "tests", "host_sim")


@pytest.fixture(scope="module")
def sim():
    so = os.path.join(SIM, "libstar_thread_host.so")
    src = os.path.join(SIM, "star_thread_host.cpp")
    deps = [src] + [os.path.join(ROOT, "mvoscalerecovery_b200", "csrc", f) for f in ("gthread.cuh", "gindex.cuh")]
    if not os.path.isfile(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-DMVOSR_THREAD_COST", "-o", so, src])
    lib = C.CDLL(so)
    lib.star_thread_run.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_float,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    return lib


def points(kind, n, rng):
    if kind == "uniform":
        u, v = rng.uniform(0, 1241, n), rng.uniform(186, 376, n)
    elif kind == "ground":
        X, Z = rng.uniform(-8, 8, n), rng.uniform(5, 40, n)
        u, v = 718.856 * X / Z + 607.19, 718.856 * 1.7 / Z + 185.2
    elif kind == "clustered":
        c = rng.uniform([100, 200], [1100, 360], (12, 2))
        k = rng.integers(0, 12, n)
        u, v = c[k, 0] + rng.normal(0, 40, n), c[k, 1] + rng.normal(0, 15, n)
    else:                                              # integer pixels: co-circular / collinear ties and duplicates everywhere
        u, v = rng.integers(0, 200, n).astype(float), rng.integers(186, 260, n).astype(float)
    return u.astype(np.float32), v.astype(np.float32)


def run(sim, u, v, cap=None, keep=None):
    n = u.shape[0]
    cap = cap or ((n + 63) // 64) * 64
    status = np.full(n, -1, np.int32); deg = np.zeros(n, np.int32); ring = np.full((n, 16), -1, np.int32); cost = np.zeros(8, np.uint64)
    k = None if keep is None else np.ascontiguousarray(keep, dtype=np.uint8)
    sim.star_thread_run(n, u.ctypes.data, v.ctypes.data, cap, 1.5, 4, 2.5, status.ctypes.data, deg.ctypes.data, ring.ctypes.data, cost.ctypes.data,
                        None if k is None else k.ctypes.data)
    return status, deg, ring, cost


def qhull_rings(u, v):
    """neighbour list of every point, counter-clockwise (image axes: x right, y down -> the kernel's 'left of p->cur' is cross > 0
    in (x, y) as stored), starting anywhere"""
    pts = np.stack([u, v], 1).astype(np.float64)
    dt = Delaunay(pts)
    indptr, idx = dt.vertex_neighbor_vertices
    hull = set(dt.convex_hull.reshape(-1).tolist())
    rings = []
    for p in range(pts.shape[0]):
        nb = idx[indptr[p]:indptr[p + 1]]
        d = pts[nb] - pts[p]
        rings.append(nb[np.argsort(np.arctan2(d[:, 1], d[:, 0]))])
    return rings, hull


@pytest.mark.parametrize("kind,n", [("uniform", 2000), ("ground", 2000), ("clustered", 2000), ("uniform", 300), ("uniform", 6000)])
def test_certified_stars_equal_qhull(sim, kind, n):
    rng = np.random.default_rng(sum(map(ord, kind)) + n)
    u, v = points(kind, n, rng)
    status, deg, ring, cost = run(sim, u, v)
    rings, hull = qhull_rings(u, v)
    ok = np.flatnonzero(status == 0)
    assert ok.size > 0.7 * n, (kind, ok.size, n)
    for p in ok:
        mine = ring[p, :deg[p]]
        ref = rings[p]
        assert p not in hull
        assert deg[p] == ref.size and set(mine.tolist()) == set(ref.tolist()), (kind, p, mine, ref)
        k = int(np.flatnonzero(ref == mine[0])[0])
        assert np.array_equal(np.roll(ref, -k), mine), (kind, p, mine, ref)
        d = np.hypot(u[ref] - u[p], v[ref] - v[p])
        assert d[k] == d.min()
    print("\n%s n=%d: certified %.1f %%, nocand %d, out %d, defer %d; %.1f evaluations and %.2f steps per star; lock-step efficiency %.2f (evals) %.2f (steps); R=%d; warp iterations per star %.1f (ideal %.1f), max M %d"
          % (kind, n, 100.0 * ok.size / n, np.count_nonzero(status == 1), np.count_nonzero(status == 2), np.count_nonzero(status == 3),
             cost[0] / cost[2], cost[1] / cost[2], cost[0] / max(cost[3], 1), cost[1] / max(cost[4], 1), cost[5], cost[6] / cost[2], cost[0] / cost[2] / 32.0, cost[7]))


def test_degenerate_sets_never_certify_wrongly(sim):
    """integer pixels: ties and duplicates -- whatever is certified must still be right (Qhull's own tie-break is not compared:
    a certified star has no tie among its decisions)"""
    rng = np.random.default_rng(5)
    u, v = points("integer", 1500, rng)
    status, deg, ring, cost = run(sim, u, v)
    pts = np.stack([u, v], 1).astype(np.float64)
    uniq, first = np.unique(pts, axis=0, return_index=True)
    ok = np.flatnonzero(status == 0)
    # every certified ring must be a set of true Delaunay neighbours of the de-duplicated set with strictly empty circles: check by brute force
    for p in ok[:200]:
        mine = ring[p, :deg[p]]
        for a, b in zip(mine, np.roll(mine, -1)):
            A, B, P0 = pts[a] - pts[p], pts[b] - pts[p], pts[first] - pts[p]
            m0 = A[0] * B[1] - A[1] * B[0]
            assert m0 > 0
            al, bl, sl = A @ A, B @ B, (P0 * P0).sum(1)
            det = m0 * sl + (A[1] * bl - al * B[1]) * P0[:, 0] + (al * B[0] - A[0] * bl) * P0[:, 1]
            inside = det < 0
            assert not inside.any(), (p, a, b)


@pytest.mark.parametrize("kind,frac", [("uniform", 0.84), ("ground", 0.72), ("clustered", 0.62)])
def test_filtered_index_serves_the_subset(sim, kind, frac):
    """Delaunay #2 runs over the survivors of the graph check on the index of Delaunay #1 FILTERED in place (filter_grid,
    csrc/frame_kernel.cuh: strips, bins and sub-cells kept, starts replaced by survivor counts) instead of a new build.  The same
    filter, restated sequentially in the host simulation: every star certified over the filtered index equals Qhull's star in the
    triangulation of the SURVIVORS, and about as many stars are certified as over an index built for them."""
    rng = np.random.default_rng(sum(map(ord, kind)))
    u, v = points(kind, 2000, rng)
    keep = rng.random(2000) < frac
    status, deg, ring, cost = run(sim, u, v, keep=keep)
    us, vs = u[keep], v[keep]
    m = us.shape[0]
    rings, hull = qhull_rings(us, vs)
    ok = np.flatnonzero(status[:m] == 0)
    for p in ok:
        mine = ring[p, :deg[p]]
        ref = rings[p]
        assert deg[p] == ref.size and set(mine.tolist()) == set(ref.tolist()), (kind, p, mine, ref)
        k = int(np.flatnonzero(ref == mine[0])[0])
        assert np.array_equal(np.roll(ref, -k), mine), (kind, p)
    fresh = run(sim, us, vs)[0]
    print("\n%s, %d of 2000 kept: certified over the filtered index %.1f %%, over a fresh one %.1f %%" % (kind, m, 100.0 * ok.size / m, 100.0 * np.count_nonzero(fresh == 0) / m))
    assert ok.size > 0.6 * m and ok.size > 0.85 * np.count_nonzero(fresh == 0)
