"""ctypes binding of oracle/exact_dt.c (exact Delaunay + exact predicates) -- TEST INFRASTRUCTURE."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "liboracle.so")
        if not os.path.isfile(path):
            build()
        L = ctypes.CDLL(path)
        fp = ctypes.POINTER(ctypes.c_float)
        L.oracle_delaunay.argtypes = [fp, fp, ctypes.c_int, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_uint8)]
        L.oracle_delaunay.restype = ctypes.c_int
        L.oracle_orient.argtypes = [fp, fp, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.oracle_incircle_raw.argtypes = [fp, fp] + [ctypes.c_int] * 4
        L.oracle_incircle_sos.argtypes = [fp, fp] + [ctypes.c_int] * 5
        _LIB = L
    return _LIB


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def delaunay_exact(points2d):
    """Exact Delaunay (symbolic tie-break, lowest-index duplicate kept) of float32 points.
    Returns (canonical (T,3) int32 triangles, dup mask (n,) bool)."""
    pts = np.asarray(points2d)
    x = np.ascontiguousarray(pts[:, 0], dtype=np.float32)
    y = np.ascontiguousarray(pts[:, 1], dtype=np.float32)
    assert np.array_equal(x.astype(pts.dtype), pts[:, 0]) and np.array_equal(y.astype(pts.dtype), pts[:, 1]), \
        "exact oracle takes float32-representable coordinates"
    n = x.shape[0]
    tri = np.zeros((max(2 * n, 1), 3), dtype=np.int32)
    dup = np.zeros(max(n, 1), dtype=np.uint8)
    T = lib().oracle_delaunay(_fp(x), _fp(y), n, tri.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                              dup.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
    if T < 0:
        raise RuntimeError("oracle_delaunay failed: %d" % T)
    return tri[:T].copy(), dup[:n].astype(bool)


def orient(x, y, a, b, c):
    return lib().oracle_orient(_fp(x), _fp(y), a, b, c)


def incircle_raw(x, y, a, b, c, d):
    return lib().oracle_incircle_raw(_fp(x), _fp(y), a, b, c, d)


def incircle_sos(x, y, a, b, c, d):
    return lib().oracle_incircle_sos(_fp(x), _fp(y), len(x), a, b, c, d)


def validate_delaunay(points2d, tri, dup=None):
    """Independent exact check that ``tri`` is A Delaunay triangulation of the (non-duplicate)
    points: every triangle non-degenerate, every edge shared by <=2 triangles, the boundary a
    single convex cycle containing every point, every interior edge locally Delaunay
    (exact in-circle <= 0, i.e. co-circular allowed), every non-duplicate point a vertex.
    Returns (ok, message, n_cocircular_edges)."""
    pts = np.asarray(points2d)
    x = np.ascontiguousarray(pts[:, 0], dtype=np.float32)
    y = np.ascontiguousarray(pts[:, 1], dtype=np.float32)
    n = x.shape[0]
    tri = np.asarray(tri, dtype=np.int64)
    if dup is None:
        dup = np.zeros(n, bool)
    used = np.zeros(n, bool)
    used[tri.reshape(-1)] = True
    if not np.array_equal(used, ~dup):
        return False, "vertex set differs from the non-duplicate points", 0
    edges = {}
    for t, (a, b, c) in enumerate(tri):
        o = orient(x, y, int(a), int(b), int(c))
        if o == 0:
            return False, "degenerate triangle %d" % t, 0
        if o < 0:
            b, c = c, b
        for (u, v, w) in ((a, b, c), (b, c, a), (c, a, b)):
            key = (min(u, v), max(u, v))
            edges.setdefault(key, []).append((int(u), int(v), int(w)))
    ncoc = 0
    boundary = {}
    for key, lst in edges.items():
        if len(lst) > 2:
            return False, "edge %s in %d triangles" % (key, len(lst)), 0
        if len(lst) == 1:
            u, v, w = lst[0]
            boundary[u] = v                       # directed CCW boundary edge
        else:
            (u, v, w), (u2, v2, w2) = lst
            if (u, v) != (v2, u2):
                return False, "edge %s not oppositely oriented" % (key,), 0
            s = incircle_raw(x, y, u, v, w, w2)
            if s > 0:
                return False, "edge %s not locally Delaunay" % (key,), 0
            ncoc += s == 0
    if boundary:
        start = next(iter(boundary))
        cyc = [start]
        cur = boundary[start]
        while cur != start:
            cyc.append(cur)
            if cur not in boundary or len(cyc) > len(boundary):
                return False, "boundary is not a single cycle", 0
            cur = boundary[cur]
        if len(cyc) != len(boundary):
            return False, "boundary has several cycles", 0
        m = len(cyc)
        for i in range(m):
            if orient(x, y, cyc[i], cyc[(i + 1) % m], cyc[(i + 2) % m]) < 0:
                return False, "boundary not convex", 0
        # Euler: T = 2V - 2 - B for a triangulated disk
        V = int(used.sum())
        if tri.shape[0] != 2 * V - 2 - m:
            return False, "triangle count %d != 2V-2-B = %d" % (tri.shape[0], 2 * V - 2 - m), 0
    return True, "ok", int(ncoc)
