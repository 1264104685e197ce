"""oracle/extras.py (CPU restatements of the steps around the hot path) pinned against outputs of the reference itself and of
OpenCV, generated in the build container: tests/golden/scripts.npz, pose.npz, compat_api.npz.  No GPU."""
import os

import numpy as np
import pytest

from oracle import extras as X

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = lambda n: np.load(os.path.join(ROOT, "tests", "golden", n))


def test_motion2pose_equals_reference_get_path():
    z = G("scripts.npz")
    got = X.motion2pose(z["gp_motions"], z["gp_scales"])
    np.testing.assert_allclose(got, z["gp_poses"], rtol=1e-13, atol=1e-13)
    assert np.array_equal(got[0], np.eye(4)[:3].reshape(-1))
    assert X.motion2pose(np.zeros((0, 12))).shape == (1, 12)


def test_recover_pose_equals_opencv():
    """decomposeEssentialMat + the four-way cheirality count against cv2.recoverPose's own R, t and mask count."""
    z = G("pose.npz")
    off = z["offsets"]
    for f in range(off.shape[0] - 1):
        a, e = off[f], off[f + 1]
        cur = np.stack([z["cur_u"][a:e], z["cur_v"][a:e]], 1); ref = np.stack([z["ref_u"][a:e], z["ref_v"][a:e]], 1)
        R, t, mask, counts = X.recover_pose(z["E"][f], cur, ref, 718.856, 718.856, 607.1928, 185.2157)
        np.testing.assert_allclose(R.reshape(-1), z["R"][f], atol=1e-9)
        np.testing.assert_allclose(t, z["t"][f], atol=1e-9)
        assert int(mask.sum()) == z["n_good"][f] == max(counts)
    R1, R2, t = X.decompose_essential(z["E"][1])
    for R in (R1, R2):
        np.testing.assert_allclose(R @ R.T, np.eye(3), atol=1e-12)
        assert abs(np.linalg.det(R) - 1) < 1e-12
    assert abs(np.linalg.norm(t) - 1) < 1e-12


def test_triangle_primitives_equal_reference():
    z = G("compat_api.npz")
    f3, f2, tri = z["f3"], z["f2"], z["tri"]
    n, h, my = X.triangle_planes(f3, tri)
    pitch = np.degrees(np.arcsin(-n[:, 1] / np.linalg.norm(n, axis=1)))
    loose, tight = pitch < -80, pitch < -85
    level = 0.9 * np.median(h[loose])
    assert level == z["height_level"] and np.array_equal(tri[tight & (h > level)].reshape(-1), z["flat_ids"])
    assert np.array_equal(h[loose], z["flat_heights"])
    flagged, incident = X.triangle_votes(tri, f2[:, 1], f3[:, 2], f3.shape[0])
    assert np.array_equal(1.0 - flagged, z["outliers"]) and incident.sum() == 3 * tri.shape[0]
    s = G("scripts.npz")
    fl, inc = X.triangle_votes(s["sc_tri"], s["frame2"][:, 1], s["sc_f3_1"][:, 2], s["sc_f3_1"].shape[0])
    assert np.array_equal((1.0 + inc - 2.0 * fl) >= 0, s["sc_find_outliers"])          # the older estimator's rule (scale_calculator.py:151-167)


def test_depth_from_mesh_is_consistent():
    from scipy.spatial import Delaunay
    z = G("compat_api.npz")
    dt = Delaunay(z["f2"])
    n, h, _ = X.triangle_planes(z["f3"], dt.simplices)
    ln = np.linalg.norm(n, axis=1); sg = np.where(n[:, 1] < 0, -1.0, 1.0)
    datas = np.hstack([n / ln[:, None] * sg[:, None], (sg / ln)[:, None]])
    depth, ids = X.depth_from_mesh(1241, 376, 718.856, 718.856, 607.1928, 185.2157, dt, datas)
    assert depth.shape == (376, 1241) and (depth[ids < 0] == 0).all() and (ids >= 0).mean() > 0.2
    # on the plane the features came from, the depth of a pixel is the depth of the ground under it (2 % depth noise in the data)
    v, u = np.nonzero(ids >= 0)
    want = 1.7 * 718.856 / (v - 185.2157)
    sel = (v > 230) & (np.abs(depth[v, u] / want - 1) < 0.5)
    assert sel.mean() > 0.3 and np.median(np.abs(depth[v, u][sel] / want[sel] - 1)) < 0.1
