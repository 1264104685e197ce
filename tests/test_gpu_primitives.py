"""Stand-alone primitives of the C ABI (mvosr_triangle_planes / _triangle_votes / _ransac_planes / _integrate_paths)
against the oracle, numpy/LAPACK and the reference's own outputs (tests/golden/compat_api.npz).

Tolerances: counts, chosen hypothesis and masks bit-exact; float64 results 1e-9 relative (closed forms vs LAPACK)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RTOL = 1e-9


def _t(engine, a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).to(engine.device)


@pytest.fixture(scope="module")
def ref():
    return np.load(os.path.join(ROOT, "tests", "golden", "compat_api.npz"))


def test_triangle_planes_vs_lapack_and_reference(engine, ref):
    """n = P^-1 1 (np.matrix(...).I of rescale.py:79), height = 1/|n|, mean Y; then the reference's flat_selection outputs."""
    f3, tri = ref["f3"], ref["tri"].astype(np.int32)
    out = engine.triangle_planes(_t(engine, tri), _t(engine, f3))
    n = out["normal"].cpu().numpy(); h = out["height"].cpu().numpy(); my = out["mean_y"].cpu().numpy()
    want = np.stack([np.linalg.inv(f3[t]) @ np.ones(3) for t in tri])
    np.testing.assert_allclose(n, want, rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(h, 1.0 / np.linalg.norm(want, axis=1), rtol=1e-8)
    np.testing.assert_allclose(my, f3[tri][:, :, 1].mean(1), rtol=1e-14)
    # the reference's own gates on top of the GPU planes (rescale.py:83-96)
    ln = np.linalg.norm(n, axis=1)
    pitch = np.degrees(np.arcsin(-n[:, 1] / ln))
    loose, tight = pitch < -80, pitch < -85
    level = 0.9 * np.median(h[loose])
    np.testing.assert_allclose(level, ref["height_level"], rtol=RTOL)
    assert np.array_equal(tri[tight & (h > level)].reshape(-1), ref["flat_ids"])
    np.testing.assert_allclose(h[loose], ref["flat_heights"], rtol=RTOL)


def test_triangle_planes_singular_is_nan(engine):
    xyz = np.array([[1.0, 1.0, 1.0], [2.0, 2.0, 2.0], [3.0, 3.0, 3.0], [0.0, 1.0, 5.0]])
    out = engine.triangle_planes(_t(engine, np.array([[0, 1, 2]], np.int32)), _t(engine, xyz))
    assert np.isnan(out["height"].cpu().numpy()[0]) and np.isnan(out["normal"].cpu().numpy()).all()


def test_triangle_votes_vs_reference(engine, ref):
    """find_outliers of rescale.py:63-72 = 1 - flagged; incident = triangles per vertex."""
    f3, f2, tri = ref["f3"], ref["f2"], ref["tri"].astype(np.int32)
    out = engine.triangle_votes(_t(engine, tri), _t(engine, f2[:, 1]), _t(engine, f3[:, 2]))
    assert np.array_equal(1.0 - out["flagged"].cpu().numpy(), ref["outliers"])
    assert np.array_equal(out["incident"].cpu().numpy(), np.bincount(tri.reshape(-1), minlength=f3.shape[0]))


def test_ransac_planes_vs_oracle(engine):
    """Same Philox stream, same sequential bookkeeping: chosen hypothesis, count and hypotheses used are exact."""
    from oracle import pipeline as P
    rng = np.random.default_rng(5)
    lists, frames = [], []
    for s, (n, out_frac, noise) in enumerate([(400, 0.0, 0.002), (900, 0.3, 0.003), (60, 0.5, 0.004), (12, 0.0, 0.001),
                                               (2500, 0.45, 0.006), (3, 0.0, 0.0), (300, 0.9, 0.01)]):
        x = rng.uniform(-8, 8, n); z = rng.uniform(5, 40, n)
        y = 1.7 + 0.02 * x - 0.01 * z + noise * rng.standard_normal(n)
        bad = rng.random(n) < out_frac
        y[bad] -= rng.uniform(0.2, 1.5, bad.sum())
        lists.append(np.stack([x, y, z], 1)); frames.append(10 + 3 * s)
    # a vertex list with multiplicity (every point three to six times, as rescale.py:101 builds it): hypotheses that draw the
    # same point twice are degenerate (the reference's SVD returns an arbitrary plane for them; here they are skipped)
    lists.append(np.repeat(lists[1][:150], rng.integers(3, 7, 150), axis=0)); frames.append(99)
    off = np.zeros(len(lists) + 1, np.int32)
    np.cumsum([a.shape[0] for a in lists], out=off[1:])
    xyz = np.concatenate(lists, 0)
    for stop, iters in ((True, 100), (False, 64), (True, 500)):
        out = engine.ransac_planes(_t(engine, off), _t(engine, xyz), iterations=iters, threshold=0.005, stop_at_goal=stop,
                                   seed=77, frame_index=_t(engine, np.asarray(frames, np.int32)), seq_id=3)
        m = out["model"].cpu().numpy(); ic = out["ic"].cpu().numpy(); bh = out["best_hyp"].cpu().numpy(); hu = out["hyps_used"].cpu().numpy()
        for s, pts in enumerate(lists):
            r = P.ransac_plane(pts, 77, frames[s], seq=3, max_iterations=iters, thr=0.005, stop_at_goal=stop)
            assert (ic[s], bh[s], hu[s]) == (r["ic"], r["best_hyp"], r["hyps_used"]), (s, stop, iters)
            want = np.asarray(r["model"]) * (1.0 if r["model"][1] >= 0 else -1.0)
            np.testing.assert_allclose(m[s], want, rtol=1e-7, atol=1e-11)
            np.testing.assert_allclose(np.linalg.norm(m[s]), 1.0, rtol=1e-12)


def test_ransac_planes_too_few_points(engine):
    off = np.array([0, 2, 2], np.int32)
    out = engine.ransac_planes(_t(engine, off), _t(engine, np.zeros((2, 3))))
    assert np.isnan(out["model"].cpu().numpy()).all() and (out["best_hyp"].cpu().numpy() == -1).all()
    assert (out["ic"].cpu().numpy() == 0).all() and (out["hyps_used"].cpu().numpy() == 0).all()


def _motion2pose(motions, scales):
    """Left-to-right restatement of get_path + motion2pose (src/main_offline.py:95-119)."""
    poses = np.zeros((motions.shape[0] + 1, 12))
    cur = np.eye(4)
    poses[0] = cur[:3].reshape(-1)
    for i in range(motions.shape[0]):
        m = np.eye(4)
        m[:3] = motions[i].reshape(3, 4)
        m[:3, 3] *= scales[i]
        cur = cur @ m
        poses[i + 1] = cur[:3].reshape(-1)
    return poses


def test_integrate_paths_vs_sequential_product(engine):
    from mvoscalerecovery_b200 import synth
    rng = np.random.default_rng(9)
    lens = [4541, 1, 0, 271, 1101]
    off = np.zeros(len(lens) + 1, np.int32)
    np.cumsum(lens, out=off[1:])
    F = int(off[-1])
    mot = np.zeros((F, 12))
    for i in range(F):
        R = synth._rodrigues(*np.deg2rad(rng.uniform(-1.5, 1.5, 3)))
        t = np.array([rng.uniform(-0.05, 0.05), rng.uniform(-0.02, 0.02), 1.0]); t /= np.linalg.norm(t)
        mot[i] = np.hstack([R, t[:, None]]).reshape(-1)
    sc = rng.uniform(0.0, 1.5, F)
    poses = engine.integrate_paths(_t(engine, off), _t(engine, mot), _t(engine, sc)).cpu().numpy()
    for s in range(len(lens)):
        want = _motion2pose(mot[off[s]:off[s + 1]], sc[off[s]:off[s + 1]])
        got = poses[off[s] + s: off[s + 1] + s + 1]
        assert got.shape == want.shape
        np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-9 * max(1.0, np.abs(want).max()))
        assert np.array_equal(got[0], np.eye(4)[:3].reshape(-1))
    # without scales: unit steps
    p1 = engine.integrate_paths(_t(engine, off[:2]), _t(engine, mot[:lens[0]])).cpu().numpy()
    np.testing.assert_allclose(p1, _motion2pose(mot[:lens[0]], np.ones(lens[0])), rtol=1e-9, atol=1e-6)


def test_depth_from_mesh_vs_find_simplex(engine, ref):
    """Reconstruct.depth_generate (reconstruct.py:91-107): per-pixel tri.find_simplex (Qhull point location) and the plane
    depth h / (n . ray).  Pixels on a mesh edge belong to either neighbour; everywhere else the triangle index is exact."""
    from scipy.spatial import Delaunay
    W, H, fx, fy, cx, cy = 1241, 376, 718.856, 718.856, 607.1928, 185.2157
    f3, f2 = ref["f3"], ref["f2"]
    dt = Delaunay(f2)
    tri = dt.simplices.astype(np.int32)
    n = np.stack([np.linalg.inv(f3[t]) @ np.ones(3) for t in tri]); ln = np.linalg.norm(n, axis=1)
    sg = np.where(n[:, 1] < 0, -1.0, 1.0)
    datas = np.hstack([n / ln[:, None] * sg[:, None], (sg / ln)[:, None]])          # rows of triangle_model (reconstruct.py:70-90)
    out = engine.depth_from_mesh(W, H, fx, fy, cx, cy, _t(engine, tri), _t(engine, f2), _t(engine, datas))
    depth = out["depth"].cpu().numpy(); ids = out["tri_id"].cpu().numpy()
    v, u = np.mgrid[0:H, 0:W]
    pix = np.stack([u.ravel(), v.ravel()], 1).astype(np.float64)
    want = dt.find_simplex(pix)
    # barycentric coordinates in the reference's simplex: strictly interior pixels must agree exactly
    T = dt.transform[np.maximum(want, 0)]
    bc = np.einsum("nij,nj->ni", T[:, :2], pix - T[:, 2])
    bary = np.hstack([bc, 1 - bc.sum(1, keepdims=True)])
    interior = (want >= 0) & (bary.min(1) > 1e-7)
    got = ids.ravel()
    assert np.array_equal(got[interior], want[interior])
    outside = dt.find_simplex(pix, tol=1e-7) < 0
    assert np.all(got[outside] == -1) and np.all(depth.ravel()[outside] == 0.0)
    assert (got != want).mean() < 2e-3                                               # only pixels on mesh edges may differ
    ok = got >= 0
    xn, yn = (pix[ok, 0] - cx) / fx, (pix[ok, 1] - cy) / fy
    d = datas[got[ok]]
    np.testing.assert_allclose(depth.ravel()[ok], d[:, 3] / (d[:, 0] * xn + d[:, 1] * yn + d[:, 2]), rtol=1e-12)
    assert interior.sum() > 0.2 * W * H


def test_recover_pose_vs_opencv(engine):
    """mvosr_recover_pose_frames against cv2.recoverPose's own R, t and mask count (tests/golden/pose.npz), on essential
    matrices from cv2.findEssentialMat (the reference's call, visual_odometry.py:129-133) and on exact ones with arbitrary
    scale / sign; then the fused scale recovery runs from the recovered poses."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "pose.npz"))
    d = {k: _t(engine, z[k]) for k in ("offsets", "cur_u", "cur_v", "ref_u", "ref_v")}
    out = engine.recover_pose_frames(d["offsets"], d["cur_u"], d["cur_v"], d["ref_u"], d["ref_v"], _t(engine, z["E"]))
    poses = out["poses"].cpu().numpy().reshape(-1, 3, 4); good = out["n_good"].cpu().numpy()
    F = poses.shape[0]
    for f in range(F):
        R, t = z["R"][f].reshape(3, 3), z["t"][f]
        np.testing.assert_allclose(poses[f][:, :3], R, atol=1e-9)
        np.testing.assert_allclose(poses[f][:, 3], t, atol=1e-9)
        assert good[f].max() == z["n_good"][f], (f, good[f], z["n_good"][f])
        assert abs(np.linalg.det(poses[f][:, :3]) - 1) < 1e-12 and abs(np.linalg.norm(poses[f][:, 3]) - 1) < 1e-12
        assert np.sort(good[f])[-2] < 0.5 * good[f].max()                      # the winner is unambiguous
    # the recovered poses drive stage 1 exactly like the poses they came from
    a = engine.triangulate_frames(d["offsets"], d["cur_u"], d["cur_v"], d["ref_u"], d["ref_v"], out["poses"])
    assert np.array_equal(a["n_out"].cpu().numpy(), z["n_good"])
