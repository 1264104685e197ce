"""oracle/five_point.py (groundwork for SURVEY N1: cv2.findEssentialMat's algorithm on the CPU): the minimal solver is exact on
noise-free data, and the RANSAC's pose and inlier set agree with OpenCV's own outputs (tests/golden/pose.npz) within the noise
of the data.  OpenCV draws from its own RNG, so nothing here can be bit-exact; the tolerances say what "agree" means."""
import os

import numpy as np

from oracle import extras as X
from oracle import five_point as FP
from mvoscalerecovery_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
K = (718.856, 718.856, 607.1928, 185.2157)


def _angle(a, b):
    return np.degrees(np.arccos(np.clip(a, -1, 1))) if b is None else np.degrees(np.arccos(np.clip((np.trace(a.T @ b) - 1) / 2, -1, 1)))


def test_minimal_solver_is_exact_on_noise_free_data():
    rng = np.random.default_rng(0)
    for _ in range(40):
        R = synth._rodrigues(*rng.uniform(-0.2, 0.2, 3))
        t = rng.standard_normal(3); t /= np.linalg.norm(t)
        P = np.stack([rng.uniform(-2, 2, 5), rng.uniform(-1, 1, 5), rng.uniform(4, 20, 5)], 1)
        x1 = P[:, :2] / P[:, 2:]
        P2 = P @ R.T + t
        x2 = P2[:, :2] / P2[:, 2:]
        Et = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]]) @ R
        Et /= np.linalg.norm(Et)
        sols = FP.five_point(x1, x2)
        assert 1 <= len(sols) <= 10
        assert min(min(np.linalg.norm(E - Et), np.linalg.norm(E + Et)) for E in sols) < 1e-8
        for E in sols:                                            # every solution satisfies the five constraints and is essential
            x1h, x2h = np.hstack([x1, np.ones((5, 1))]), np.hstack([x2, np.ones((5, 1))])
            assert np.abs(np.sum(x2h * (x1h @ E.T), 1)).max() < 1e-9
            s = np.linalg.svd(E, compute_uv=False)
            assert abs(s[0] - s[1]) < 1e-7 and s[2] < 1e-7


def test_ransac_agrees_with_opencv_within_the_noise():
    z = np.load(os.path.join(ROOT, "tests", "golden", "pose.npz"))
    off = z["offsets"]
    for f in (0, 2, 4, 6):                                       # the frames whose E came from cv2.findEssentialMat
        a, e = off[f], off[f + 1]
        cur = np.stack([z["cur_u"][a:e], z["cur_v"][a:e]], 1).astype(np.float64)
        ref = np.stack([z["ref_u"][a:e], z["ref_v"][a:e]], 1).astype(np.float64)
        E, mask = FP.find_essential_ransac(cur, ref, *K, threshold=0.5, seed=f)
        R, t, _, counts = X.recover_pose(E, cur, ref, *K)
        Rcv, tcv = z["R"][f].reshape(3, 3), z["t"][f]
        assert _angle(R, Rcv) < 0.2 and _angle(float(t @ tcv), None) < 1.5          # degrees; both within the noise of the truth
        Pt = z["true_poses"][f].reshape(3, 4)
        tt = Pt[:, 3] / np.linalg.norm(Pt[:, 3])
        assert _angle(R, Pt[:, :3]) < 0.2 and _angle(float(t @ tt), None) < 1.5
        assert mask.mean() > 0.9 and max(counts) > 0.9 * (e - a)
        # OpenCV's own E scores (almost) the same inlier set under the same Sampson threshold
        x1 = np.stack([(cur[:, 0] - K[2]) / K[0], (cur[:, 1] - K[3]) / K[1]], 1)
        x2 = np.stack([(ref[:, 0] - K[2]) / K[0], (ref[:, 1] - K[3]) / K[1]], 1)
        cv_mask = FP.sampson_sq(z["E"][f].reshape(3, 3) / np.linalg.norm(z["E"][f]), x1, x2) < (0.5 / K[0]) ** 2
        assert (cv_mask != mask).mean() < 0.05


def test_device_style_solver_matches_the_lapack_one():
    """oracle/five_point_plan.py (Gauss-Jordan null space, interpolated constraints, eigenvalues by QR, 6x5 back-substitution)
    against oracle/five_point.py (SVD + symbolic expansion + eigenvectors): found solutions agree and (almost) all are found."""
    from oracle import five_point_plan as PL
    rng = np.random.default_rng(3)
    matched = total = true_found = n_true = 0
    for trial in range(80):
        R = synth._rodrigues(*rng.uniform(-0.2, 0.2, 3))
        t = rng.standard_normal(3); t /= np.linalg.norm(t)
        P = np.stack([rng.uniform(-2, 2, 5), rng.uniform(-1, 1, 5), rng.uniform(4, 20, 5)], 1)
        x1 = P[:, :2] / P[:, 2:] + 1e-3 * rng.standard_normal((5, 2)) * (trial % 2)
        P2 = P @ R.T + t
        x2 = P2[:, :2] / P2[:, 2:]
        a, b = FP.five_point(x1, x2), PL.five_point_device_style(x1, x2)
        for E in a:
            total += 1
            matched += min([min(np.linalg.norm(E - F), np.linalg.norm(E + F)) for F in b] or [9]) < 1e-6
        for F in b:                                               # nothing spurious: every device-style solution is (close to) a LAPACK one
            assert min(min(np.linalg.norm(E - F), np.linalg.norm(E + F)) for E in a) < 1e-3
        if trial % 2 == 0:
            n_true += 1
            Et = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]]) @ R
            Et /= np.linalg.norm(Et)
            true_found += min([min(np.linalg.norm(E - Et), np.linalg.norm(E + Et)) for E in b] or [9]) < 1e-7
    assert matched >= 0.98 * total and true_found >= 0.97 * n_true
