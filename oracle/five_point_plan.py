"""The five-point solver restated with DEVICE-FRIENDLY numerics only -- TEST INFRASTRUCTURE (oracle side), a blueprint for the
CUDA kernel of SURVEY N1 (mvoscalerecovery_b200/csrc/five_point.cuh + five_point_kernel.cuh, written independently in C++),
and the RANSAC around it on the shared Philox stream (``find_essential_philox``): the checker of mvosr_find_essential_frames.

``oracle/five_point.py`` leans on LAPACK (SVD for the null space, a general eigen-decomposition for the action matrix), neither
of which exists inside a kernel.  This module solves the same minimal problem with what one thread or one warp can do in
registers, and is checked against the LAPACK version in tests/test_oracle_five_point.py:

  1. null space of the 5x9 epipolar system by Gauss-Jordan elimination with full pivoting (4 free columns), orthonormalised
     by modified Gram-Schmidt;
  2. the ten cubic constraints by polynomial interpolation: they are cubics in (x, y, z) with 20 coefficients, so their values at
     20 fixed generic sample points times a constant inverse Vandermonde matrix give the 10x20 coefficient matrix -- on the GPU
     one lane per sample point evaluates det E and 2 E E^T E - tr(E E^T) E numerically, no symbolic expansion;
  3. Gauss-Jordan on the 10x20 matrix (partial pivoting), action matrix of "multiply by x" on the basis
     [x^2, xy, xz, y^2, yz, z^2, x, y, z, 1];
  4. its real eigenvalues by balancing + Hessenberg reduction + double-shift QR (here: LAPACK's driver, which is that
     sequence; in the kernel: its own balanc / elmhes / hqr); no polishing -- a Rayleigh-quotient iteration on the matrix
     changed nothing measurable once the eigenvalues came from QR and was dropped on both sides.
     (A first version formed the characteristic polynomial by Faddeev-LeVerrier and isolated its roots with a Sturm chain:
     the COEFFICIENTS are ill-conditioned when the eigenvalues spread over orders of magnitude, and 5-7 % of the solutions
     were lost; the eigenvalues of A themselves are well conditioned.);
  5. for every real root x: the remaining unknowns (y^2, yz, z^2, y, z) from the first six rows of (A - x I) v = 0 by a 6x5
     least-squares solve (normal equations, Gaussian elimination).

Status (tests/test_oracle_five_point.py, tests/test_five_point_host_sim.py): solutions agree with the LAPACK version to
< 1e-6 and > 99 % of them are found (the rest: near-double roots that one side sees as a complex pair).
"""
from __future__ import annotations

import numpy as np

from .five_point import _MONO

# 20 sample points for the interpolation: the principal lattice of degree 3 on the nodes t (i + j + k <= 3), which is unisolvent
# for cubics in three variables, and the inverse Vandermonde matrix in the monomial order (a constant table on the device)
_T = np.array([-1.0, -0.4, 0.4, 1.0])
_PTS = np.array([(_T[i], _T[j], _T[k]) for i in range(4) for j in range(4) for k in range(4) if i + j + k <= 3])
_VAND = np.array([[p[0] ** i * p[1] ** j * p[2] ** k for (i, j, k) in _MONO] for p in _PTS])
_VINV = np.linalg.inv(_VAND)


def _null4(Q):
    """Four null vectors of the 5x9 matrix Q by Gauss-Jordan with full pivoting."""
    M = Q.astype(np.float64).copy()
    cols = list(range(9))
    for r in range(5):
        sub = np.abs(M[r:, r:])
        i, j = np.unravel_index(int(np.argmax(sub)), sub.shape)
        if not sub[i, j] > 1e-300:
            return None                                            # rank-deficient sample (repeated correspondences)
        M[[r, r + i]] = M[[r + i, r]]
        M[:, [r, r + j]] = M[:, [r + j, r]]
        cols[r], cols[r + j] = cols[r + j], cols[r]
        M[r] /= M[r, r]
        for k in range(5):
            if k != r:
                M[k] -= M[k, r] * M[r]
    out = np.zeros((4, 9))
    for f in range(4):                                             # free column 5 + f set to 1
        v = np.zeros(9)
        v[5 + f] = 1.0
        v[:5] = -M[:, 5 + f]
        out[f, cols] = v
    for f in range(4):                                             # modified Gram-Schmidt: an orthonormal basis keeps the unknowns
        for g in range(f):                                         # x, y, z (coordinates relative to the fourth vector) moderate
            out[f] -= (out[f] @ out[g]) * out[g]
        out[f] /= np.linalg.norm(out[f])
    return out


def _constraints_at(E):
    """det E and the nine entries of 2 E E^T E - tr(E E^T) E, numerically, for one 3x3 matrix."""
    EEt = E @ E.T
    return np.concatenate([[np.linalg.det(E)], (2 * EEt @ E - np.trace(EEt) * E).reshape(-1)])


def _real_eigenvalues(A):
    """Real eigenvalues of the action matrix, ascending.  LAPACK's general driver (balancing, Hessenberg reduction, double-shift
    QR) stands in for the kernel's own balanc / elmhes / hqr sequence in csrc/five_point.cuh: same algorithm family, written
    independently, so agreement is to rounding, not bit for bit.  A close complex pair counts once."""
    w = np.linalg.eigvals(A)
    keep = (np.abs(w.imag) <= 1e-9 * np.maximum(1.0, np.abs(w.real))) & (w.imag >= 0)
    return sorted(w.real[keep])


def five_point_device_style(x1, x2):
    """Same contract as oracle.five_point.five_point, device-friendly numerics (see the module docstring)."""
    x1h = np.hstack([np.asarray(x1, dtype=np.float64), np.ones((5, 1))])
    x2h = np.hstack([np.asarray(x2, dtype=np.float64), np.ones((5, 1))])
    Q = np.stack([np.kron(x2h[i], x1h[i]) for i in range(5)])
    basis = _null4(Q)
    if basis is None:
        return []
    X, Y, Z, W = (b.reshape(3, 3) for b in basis)
    vals = np.stack([_constraints_at(p[0] * X + p[1] * Y + p[2] * Z + W) for p in _PTS])     # 20 x 10
    M = (_VINV @ vals).T                                                                       # 10 x 20 coefficients
    mx = np.abs(M).max(1, keepdims=True)
    if not (np.all(mx > 0) and np.all(mx < 1e300)):
        return []
    M = M / mx
    for col in range(10):
        piv = col + int(np.argmax(np.abs(M[col:, col])))
        if abs(M[piv, col]) < 1e-13:
            return []
        M[[col, piv]] = M[[piv, col]]
        M[col] /= M[col, col]
        for r in range(10):
            if r != col:
                M[r] -= M[r, col] * M[col]
    B = M[:, 10:]
    A = np.zeros((10, 10))
    A[0:6] = -B[0:6]
    A[6, 0] = A[7, 1] = A[8, 2] = A[9, 6] = 1.0
    sols = []
    if not np.isfinite(A).all():
        return []
    for x in _real_eigenvalues(A):
        # (A - x I) v = 0 with v = [x^2, xy, xz, y^2, yz, z^2, x, y, z, 1]; x known -> unknowns u = [xy, xz, y^2, yz, z^2, y, z]
        # rows 7, 8 give xy = x*y, xz = x*z; rows 0..5 are linear in (y^2, yz, z^2, y, z) once those are substituted
        R = A[0:6] - x * np.eye(10)[0:6]
        # columns: 0 x^2 (known), 1 xy = x y, 2 xz = x z, 3 y^2, 4 yz, 5 z^2, 6 x (known), 7 y, 8 z, 9 1
        L = np.stack([R[:, 3], R[:, 4], R[:, 5], R[:, 7] + x * R[:, 1], R[:, 8] + x * R[:, 2]], 1)
        rhs = -(R[:, 0] * x * x + R[:, 6] * x + R[:, 9])
        try:
            u = np.linalg.solve(L.T @ L, L.T @ rhs)               # 5x5 normal equations (Gaussian elimination on the device)
        except np.linalg.LinAlgError:
            continue
        y, z = u[3], u[4]
        Ek = x * X + y * Y + z * Z + W
        Ek = Ek / np.linalg.norm(Ek)
        # accept only what IS an essential matrix (an inaccurate eigenvalue of a near-multiple root, or an ill-conditioned 5x5 solve,
        # gives a matrix of the null space that violates the cubic constraints): nine multiply-adds per entry on the device
        EEt = Ek @ Ek.T
        if np.abs(2 * EEt @ Ek - np.trace(EEt) * Ek).max() < 1e-6 and abs(np.linalg.det(Ek)) < 1e-6:
            sols.append(Ek)
    return sols


# ---------------------------------------------------------------------------------------------------------------------------
# RANSAC on a DEFINED sample stream (OpenCV's own RNG cannot be reproduced): the checker of mvosr_find_essential_frames.
#   key = (seed lo, seed hi); (r0..r3) = Philox4x32-10(counter = (hyp, frame, seq, 1)); r4 = word 0 of counter (hyp, frame, seq, 2)
#   p_k = (r_k * (n - k)) >> 32, then + 1 for every earlier position (ascending) it is >= to   -> five distinct positions
# Selection: the candidate with the most Sampson inliers; ties go to the lowest (hypothesis, candidate) pair, candidates in
# ascending order of their eigenvalue.  `hypotheses` is the maximum (OpenCV's maxIters); confidence > 0 (OpenCV's prob) stops
# after the first ROUND of 128 hypotheses at whose end  tried >= log(1 - confidence) / log(1 - w^5),  w = best / n.
ROUND = 128


def enough_hypotheses(tried, best_count, n, confidence):
    if not confidence > 0.0 or best_count <= 0 or confidence >= 1.0:
        return False
    w5 = (best_count / n) ** 5
    if w5 >= 1.0:
        return True
    return tried * np.log(1.0 - w5) <= np.log(1.0 - confidence)
def sample5_positions(seed, hyp, frame, seq, n):
    from .philox import MASK, philox4x32_10
    key = (seed & MASK, (seed >> 32) & MASK)
    r = list(philox4x32_10((hyp, frame, seq, 1), key)) + [philox4x32_10((hyp, frame, seq, 2), key)[0]]
    chosen = []
    for k in range(5):
        p = (r[k] * (n - k)) >> 32
        for a in sorted(chosen):
            if p >= a:
                p += 1
        chosen.append(p)
    return chosen


def sampson_inlier(E, x1, x2, thr2):
    """s^2 < thr2 * (|E x1|_xy^2 + |E^T x2|_xy^2), the division-free form the kernel uses."""
    x1h = np.hstack([x1, np.ones((x1.shape[0], 1))])
    x2h = np.hstack([x2, np.ones((x2.shape[0], 1))])
    Ex1 = x1h @ E.T
    Etx2 = x2h @ E
    s = np.sum(x2h * Ex1, 1)
    return s * s < thr2 * (Ex1[:, 0] ** 2 + Ex1[:, 1] ** 2 + Etx2[:, 0] ** 2 + Etx2[:, 1] ** 2)


def find_essential_philox(px_cur, px_ref, fx, fy, cx, cy, hypotheses=128, threshold=0.5, seed=0, frame=0, seq=0, solver=None, confidence=0.0):
    """(E (3,3), mask (n,) bool, n_inliers, best_hyp, hyps_used) for one frame; px_* (n,2) pixels (float32 values).  x_ref^T E x_cur = 0 in
    normalised coordinates, cv2.findEssentialMat(px_cur, px_ref, K)'s convention (visual_odometry.py:129-130)."""
    solver = solver or five_point_device_style
    px_cur = np.asarray(px_cur, dtype=np.float64)
    px_ref = np.asarray(px_ref, dtype=np.float64)
    x1 = np.stack([(px_cur[:, 0] - cx) / fx, (px_cur[:, 1] - cy) / fy], 1)
    x2 = np.stack([(px_ref[:, 0] - cx) / fx, (px_ref[:, 1] - cy) / fy], 1)
    n = x1.shape[0]
    thr2 = (threshold / (0.5 * (fx + fy))) ** 2
    best = (0, None, -1)
    used = 0
    if n >= 5:
        for hyp in range(hypotheses):
            if hyp > 0 and hyp % ROUND == 0 and enough_hypotheses(hyp, best[0], n, confidence):
                break
            used = hyp + 1
            idx = sample5_positions(seed, hyp, frame, seq, n)
            for E in solver(x1[idx], x2[idx]):
                cnt = int(sampson_inlier(E, x1, x2, thr2).sum())
                if cnt > best[0]:
                    best = (cnt, E, hyp)
    if best[1] is None:
        return np.zeros((3, 3)), np.zeros(n, bool), 0, -1, used
    return best[1], sampson_inlier(best[1], x1, x2, thr2), best[0], best[2], used
