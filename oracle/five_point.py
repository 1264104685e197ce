"""Five-point essential matrix + RANSAC on the CPU with LAPACK -- TEST INFRASTRUCTURE (oracle side) for SURVEY N1.

Part of ``oracle/``: imported only by ``tests/``.  No product code uses it: the product's own five-point RANSAC is the CUDA
kernel behind mvosr_find_essential_frames (csrc/five_point.cuh, five_point_kernel.cuh); this module is the independent
LAPACK-based statement its minimal solver is checked against (tests/test_five_point_host_sim.py), and
oracle/five_point_plan.py is the restatement on the shared Philox stream that checks the RANSAC around it.

What it restates: ``cv2.findEssentialMat(px_cur, px_ref, cameraMatrix=K, method=cv2.RANSAC, prob=0.999, threshold=0.5)`` as
the reference calls it (src/thirdparty/MonocularVO/visual_odometry.py:129-130).  The algorithm lives in OpenCV (4.13 in this
image, calib3d five-point.cpp + the generic RANSAC of ptsetreg.cpp), not in the reference tree: the minimal solver is Nister's
five-point problem (here in Stewenius' Groebner-basis / action-matrix form: null space of the 5x9 epipolar system, the ten
cubic constraints det E = 0 and 2 E E^T E - tr(E E^T) E = 0, Gauss-Jordan, eigenvectors of the 10x10 multiplication matrix),
hypotheses are scored by the Sampson distance against ``threshold`` (in pixels, divided by the focal length), and the number
of iterations adapts to the inlier ratio for the requested confidence.

Parity: OpenCV draws its samples from its own RNG, so the chosen hypothesis cannot be reproduced; what is pinned
(tests/test_oracle_five_point.py) is (i) exactness of the minimal solver on noise-free data, (ii) agreement of the recovered
pose and of the inlier set with OpenCV's own outputs (tests/golden/pose.npz) within the noise of the data.  "Parity unpinned"
at the bit level, by construction.
"""
from __future__ import annotations

import numpy as np

# monomials of degree <= 3 in (x, y, z), Stewenius' order: the ten cubics first, then the basis of the quotient ring
_MONO = [(3, 0, 0), (2, 1, 0), (2, 0, 1), (1, 2, 0), (1, 1, 1), (1, 0, 2), (0, 3, 0), (0, 2, 1), (0, 1, 2), (0, 0, 3),
         (2, 0, 0), (1, 1, 0), (1, 0, 1), (0, 2, 0), (0, 1, 1), (0, 0, 2), (1, 0, 0), (0, 1, 0), (0, 0, 1), (0, 0, 0)]


def _pmul(a, b):
    """Product of two polynomials stored as (4,4,4) coefficient grids (exponents of x, y, z)."""
    out = np.zeros((4, 4, 4))
    for i, j, k in zip(*np.nonzero(a)):
        sub = b[:4 - i, :4 - j, :4 - k]
        out[i:, j:, k:] += a[i, j, k] * sub
    return out


def _lin(cx, cy, cz, c1):
    p = np.zeros((4, 4, 4))
    p[1, 0, 0], p[0, 1, 0], p[0, 0, 1], p[0, 0, 0] = cx, cy, cz, c1
    return p


def five_point(x1, x2):
    """All real essential matrices E with x2_i^T E x1_i = 0 for five correspondences in normalised image coordinates.
    x1, x2: (5,2).  Returns a list of 3x3 arrays with unit Frobenius norm (up to ten)."""
    x1h = np.hstack([np.asarray(x1, dtype=np.float64), np.ones((5, 1))])
    x2h = np.hstack([np.asarray(x2, dtype=np.float64), np.ones((5, 1))])
    Q = np.stack([np.kron(x2h[i], x1h[i]) for i in range(5)])                    # row-major vec(E)
    basis = np.linalg.svd(Q)[2][5:9]                                               # null space: E = x X + y Y + z Z + W
    X, Y, Z, W = (b.reshape(3, 3) for b in basis)
    E = [[_lin(X[r, c], Y[r, c], Z[r, c], W[r, c]) for c in range(3)] for r in range(3)]
    # det E
    det = (_pmul(_pmul(E[0][0], E[1][1]), E[2][2]) + _pmul(_pmul(E[0][1], E[1][2]), E[2][0]) + _pmul(_pmul(E[0][2], E[1][0]), E[2][1])
           - _pmul(_pmul(E[0][2], E[1][1]), E[2][0]) - _pmul(_pmul(E[0][1], E[1][0]), E[2][2]) - _pmul(_pmul(E[0][0], E[1][2]), E[2][1]))
    # E E^T, its trace, and 2 E E^T E - tr(E E^T) E
    EEt = [[sum(_pmul(E[r][k], E[c][k]) for k in range(3)) for c in range(3)] for r in range(3)]
    tr = EEt[0][0] + EEt[1][1] + EEt[2][2]
    cons = [det]
    for r in range(3):
        for c in range(3):
            cons.append(2 * sum(_pmul(EEt[r][k], E[k][c]) for k in range(3)) - _pmul(tr, E[r][c]))
    M = np.array([[p[m] for m in _MONO] for p in cons])                            # 10 x 20
    # Gauss-Jordan with partial pivoting on the ten cubic columns: M -> [I | B]
    M = M.copy()
    for col in range(10):
        piv = col + int(np.argmax(np.abs(M[col:, col])))
        if abs(M[piv, col]) < 1e-14:
            return []
        M[[col, piv]] = M[[piv, col]]
        M[col] /= M[col, col]
        for r in range(10):
            if r != col:
                M[r] -= M[r, col] * M[col]
    B = M[:, 10:]
    # multiplication by x in the basis [x^2, xy, xz, y^2, yz, z^2, x, y, z, 1]
    A = np.zeros((10, 10))
    A[0:6] = -B[[0, 1, 2, 3, 4, 5]]
    A[6, 0] = A[7, 1] = A[8, 2] = A[9, 6] = 1.0
    w, V = np.linalg.eig(A)
    sols = []
    for k in range(10):
        if abs(w[k].imag) > 1e-9 * max(1.0, abs(w[k])) or abs(V[9, k]) < 1e-14:
            continue
        v = (V[:, k] / V[9, k]).real
        Ek = v[6] * X + v[7] * Y + v[8] * Z + W
        sols.append(Ek / np.linalg.norm(Ek))
    return sols


def sampson_sq(E, x1, x2):
    """Squared Sampson distance of every correspondence (normalised coordinates) to the epipolar constraint of E."""
    x1h = np.hstack([x1, np.ones((x1.shape[0], 1))])
    x2h = np.hstack([x2, np.ones((x2.shape[0], 1))])
    Ex1 = x1h @ E.T
    Etx2 = x2h @ E
    num = np.sum(x2h * Ex1, 1) ** 2
    return num / (Ex1[:, 0] ** 2 + Ex1[:, 1] ** 2 + Etx2[:, 0] ** 2 + Etx2[:, 1] ** 2)


def find_essential_ransac(px1, px2, fx, fy, cx, cy, threshold=0.5, prob=0.999, max_iterations=1000, seed=0):
    """RANSAC over five-point hypotheses.  px1 = px_cur, px2 = px_ref (pixels).  Returns (E, inlier mask): x2^T E x1 = 0 in
    normalised coordinates, the convention of cv2.findEssentialMat(px_cur, px_ref, K)."""
    x1 = np.stack([(px1[:, 0] - cx) / fx, (px1[:, 1] - cy) / fy], 1).astype(np.float64)
    x2 = np.stack([(px2[:, 0] - cx) / fx, (px2[:, 1] - cy) / fy], 1).astype(np.float64)
    n = x1.shape[0]
    thr2 = (threshold / (0.5 * (fx + fy))) ** 2
    rng = np.random.default_rng(seed)
    best_E, best_mask, best_cnt = None, np.zeros(n, bool), 0
    it, need = 0, max_iterations
    while it < need:
        it += 1
        idx = rng.choice(n, 5, replace=False)
        for E in five_point(x1[idx], x2[idx]):
            mask = sampson_sq(E, x1, x2) < thr2
            cnt = int(mask.sum())
            if cnt > best_cnt:
                best_E, best_mask, best_cnt = E, mask, cnt
                w = cnt / n
                need = min(max_iterations, int(np.ceil(np.log(1 - prob) / np.log(max(1 - w ** 5, 1e-12))))) if w < 1 else it
    return best_E, best_mask
