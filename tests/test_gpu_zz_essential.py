"""mvosr_find_essential_frames (five-point RANSAC on the Philox stream; SURVEY N1, first half) on the GPU, through the C ABI.

What is compared, and how strictly:
  * with itself -- hard: the mask is the Sampson test of the returned matrix, the count is the mask's, the matrix is a unit-norm
    essential matrix, edge frames (fewer than five correspondences, a rank-deficient frame) report "no model";
  * with the host evaluation of the same arithmetic (tests/host_sim: csrc/five_point.cuh compiled by g++) -- BIT FOR BIT in every
    frame: the five-point translation unit is compiled with -fmad=false, so every FP64 operation is individually rounded and the
    winning hypothesis, the inlier mask and the essential matrix are reproducible on any IEEE-754 machine;
  * with the oracle's golden (tests/golden/essential.npz, written by the independent LAPACK-based Python restatement) -- the
    inlier count within five correspondences (of ~400), the same winner in most frames and then the same matrix to 1e-6 and the
    same mask: the two solvers find the same roots to rounding, not to the bit (LAPACK's eigenvalue driver against the kernel's
    own balancing + Hessenberg + double-shift QR), which can turn a near-double real root into a complex pair and flip a
    correspondence that sits on the threshold, moving the winner between hypotheses of equal support;
  * with the truth -- the true matches recovered, the mismatches rejected, and the pose behind the matrix within the noise
    (the tolerances of tests/test_oracle_five_point.py against OpenCV's own output).
It runs last among the GPU tests on purpose (file name): it is the newest kernel."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
K = (718.856, 718.856, 607.1928, 185.2157)


def _t(engine, a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).to(engine.device)


def _norm(z, a, e):
    x1 = np.stack([(z["cur_u"][a:e].astype(np.float64) - K[2]) / K[0], (z["cur_v"][a:e].astype(np.float64) - K[3]) / K[1]], 1)
    x2 = np.stack([(z["ref_u"][a:e].astype(np.float64) - K[2]) / K[0], (z["ref_v"][a:e].astype(np.float64) - K[3]) / K[1]], 1)
    return x1, x2


def _sampson_terms(E, x1, x2):
    x1h = np.hstack([x1, np.ones((x1.shape[0], 1))]); x2h = np.hstack([x2, np.ones((x2.shape[0], 1))])
    Ex1 = x1h @ E.T; Etx2 = x2h @ E
    s = np.sum(x2h * Ex1, 1)
    return s * s, Ex1[:, 0] ** 2 + Ex1[:, 1] ** 2 + Etx2[:, 0] ** 2 + Etx2[:, 1] ** 2


def _angle_deg(c):
    return np.degrees(np.arccos(np.clip(c, -1, 1)))


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "essential.npz"))


@pytest.fixture(scope="module")
def run(engine, golden):
    z = golden
    d = {k: _t(engine, z[k]) for k in ("offsets", "cur_u", "cur_v", "ref_u", "ref_v")}
    out = engine.find_essential_frames(d["offsets"], d["cur_u"], d["cur_v"], d["ref_u"], d["ref_v"], hypotheses=int(z["hypotheses"]),
                                       threshold=float(z["threshold"]), seed=int(z["seed"]), seq_id=int(z["seq"]))
    import torch
    torch.cuda.synchronize()
    return d, {k: v.cpu().numpy() for k, v in out.items()}


def test_self_consistency_and_edge_frames(golden, run):
    z, (_, out) = golden, run
    off = z["offsets"]
    thr2 = (float(z["threshold"]) / (0.5 * (K[0] + K[1]))) ** 2
    for f in range(len(off) - 1):
        a, e = off[f], off[f + 1]
        E = out["essential"][f].reshape(3, 3); mask = out["e_mask"][a:e].astype(bool)
        assert out["n_inliers"][f] == mask.sum()
        if out["best_hyp"][f] < 0:
            assert not E.any() and not mask.any()
            continue
        assert 0 <= out["best_hyp"][f] < int(z["hypotheses"])
        assert abs(np.linalg.norm(E) - 1) < 1e-12
        s = np.linalg.svd(E, compute_uv=False)
        assert abs(s[0] - s[1]) < 1e-5 and s[2] < 1e-5
        num, den = _sampson_terms(E, *_norm(z, a, e))
        clear = np.abs(num - thr2 * den) > 1e-9 * thr2 * den                          # away from the threshold the mask is determined
        assert np.array_equal(mask[clear], (num < thr2 * den)[clear])
    assert out["best_hyp"][8] == -1 and out["best_hyp"][10] == -1                    # four correspondences; one correspondence six times
    assert out["best_hyp"][9] >= 0 and out["n_inliers"][9] == 5                      # exactly the minimal sample: all five fit


def test_against_the_oracle_golden_and_the_truth(golden, run):
    z, (_, out) = golden, run
    off = z["offsets"]
    same = 0
    for f in range(8):
        a, e = off[f], off[f + 1]
        E = out["essential"][f].reshape(3, 3); mask = out["e_mask"][a:e].astype(bool)
        assert abs(int(out["n_inliers"][f]) - int(z["n_inliers"][f])) <= 5, (f, out["n_inliers"][f], z["n_inliers"][f])
        truth = z["true_match"][a:e]
        assert (mask & truth).sum() >= 0.97 * truth.sum() and (mask & ~truth).sum() <= 4
        if out["best_hyp"][f] == z["best_hyp"][f]:
            same += 1
            Eo = z["E"][f].reshape(3, 3)
            assert min(np.abs(E - Eo).max(), np.abs(E + Eo).max()) < 1e-6
            assert (mask != z["mask"][a:e].astype(bool)).sum() <= 1
    assert same >= 5, same


def test_bit_identical_to_the_host_evaluation_of_the_same_arithmetic(golden, run):
    """The five-point translation unit is compiled with -fmad=false (csrc/five_point_api.cu): every FP64 operation of the solver
    and of the Sampson test is individually rounded, so the device must reproduce, BIT FOR BIT, what the same sequence of
    IEEE-754 operations gives on the host (tests/host_sim compiles csrc/five_point.cuh with g++, no FMA contraction on x86-64
    either): the same winning hypothesis, the same inlier count, an identical mask and an identical essential matrix in every
    frame.  This checks the device compilation and the kernel's control flow (staging, reductions, winner hand-over), not the
    algorithm -- that is test_against_the_oracle_golden_and_the_truth (independent LAPACK-based Python solver) and
    tests/test_five_point_host_sim.py::test_kernel_numerics_agree_with_opencv_own_output."""
    from test_five_point_host_sim import _ransac, load_host_sim
    L = load_host_sim()
    z, (_, out) = golden, run
    off = z["offsets"]
    for f in range(len(off) - 1):
        a, e = off[f], off[f + 1]
        E, mask, cnt, hyp = _ransac(L, *(z[k][a:e] for k in ("cur_u", "cur_v", "ref_u", "ref_v")), int(z["hypotheses"]), float(z["threshold"]),
                                    int(z["seed"]), f, int(z["seq"]))
        assert (int(out["n_inliers"][f]), int(out["best_hyp"][f])) == (cnt, hyp), (f, out["n_inliers"][f], out["best_hyp"][f], cnt, hyp)
        assert np.array_equal(out["e_mask"][a:e].astype(bool), mask), f
        assert np.array_equal(out["essential"][f].reshape(3, 3), E), (f, np.abs(out["essential"][f].reshape(3, 3) - E).max())


def test_pose_from_the_gpu_essential_matrix(engine, golden, run):
    """find_essential -> recover_pose -> triangulation: the chain of visual_odometry.py:129-147 without OpenCV."""
    z, (d, out) = golden, run
    ess, emask = _t(engine, out["essential"][:8]), _t(engine, out["e_mask"])
    off8 = _t(engine, z["offsets"][:9])                               # the eight real frames (the edge frames have no model)
    pose = engine.recover_pose_frames(off8, d["cur_u"], d["cur_v"], d["ref_u"], d["ref_v"], ess, e_mask=emask)
    poses = pose["poses"].cpu().numpy().reshape(-1, 3, 4); good = pose["n_good"].cpu().numpy()
    for f in range(8):
        P = z["true_poses"][f].reshape(3, 4)
        tt = P[:, 3] / np.linalg.norm(P[:, 3])
        assert _angle_deg((np.trace(poses[f][:, :3].T @ P[:, :3]) - 1) / 2) < 0.2
        assert _angle_deg(float(poses[f][:, 3] @ tt)) < 1.5
        assert good[f].max() >= 0.97 * out["n_inliers"][f]
    tri = engine.triangulate_frames(off8, d["cur_u"], d["cur_v"], d["ref_u"], d["ref_v"], pose["poses"], e_mask=emask)
    n_out = tri["n_out"].cpu().numpy()
    assert np.array_equal(n_out, good.max(1))


def test_hypothesis_count_and_frame_index(engine, golden, run):
    """More hypotheses never lower the support; frame_index re-addresses the stream (a shard computes what the whole does)."""
    z, (d, out) = golden, run
    import torch
    more = engine.find_essential_frames(d["offsets"], d["cur_u"], d["cur_v"], d["ref_u"], d["ref_v"], hypotheses=200, threshold=float(z["threshold"]),
                                        seed=int(z["seed"]), seq_id=int(z["seq"]), confidence=0.0)
    assert (more["n_inliers"].cpu().numpy() >= out["n_inliers"]).all()
    n = np.diff(z["offsets"])
    assert np.array_equal(more["hyps_used"].cpu().numpy(), np.where(n >= 5, 200, 0))
    assert np.array_equal(out["hyps_used"], np.where(n >= 5, int(z["hypotheses"]), 0))
    # OpenCV's adaptive count (prob = 0.999, maxIters = 1000: the reference's call), checked per round of 128: ~80 % inliers stop
    # after the first round, the frame without a model runs to the end; the first round is the same stream as above
    ada = engine.find_essential_frames(d["offsets"], d["cur_u"], d["cur_v"], d["ref_u"], d["ref_v"], hypotheses=1000, threshold=float(z["threshold"]),
                                       seed=int(z["seed"]), seq_id=int(z["seq"]), confidence=0.999)
    used = ada["hyps_used"].cpu().numpy()
    assert (used[:8] == 128).all() and used[8] == 0 and used[9] == 128 and used[10] == 1000
    assert (ada["n_inliers"].cpu().numpy() >= out["n_inliers"]).all() and (ada["n_inliers"].cpu().numpy() <= more["n_inliers"].cpu().numpy()).all()
    off = z["offsets"]
    a, e = int(off[3]), int(off[6])                                   # frames 3..5 as their own batch, addressed as frames 3..5
    sub_off = _t(engine, (off[3:7] - off[3]).astype(np.int32))
    fi = _t(engine, np.arange(3, 6, dtype=np.int32))
    sub = engine.find_essential_frames(sub_off, d["cur_u"][a:e].contiguous(), d["cur_v"][a:e].contiguous(), d["ref_u"][a:e].contiguous(),
                                       d["ref_v"][a:e].contiguous(), hypotheses=int(z["hypotheses"]), threshold=float(z["threshold"]),
                                       seed=int(z["seed"]), frame_index=fi, seq_id=int(z["seq"]))
    torch.cuda.synchronize()
    assert np.array_equal(sub["essential"].cpu().numpy(), out["essential"][3:6])
    assert np.array_equal(sub["e_mask"].cpu().numpy(), out["e_mask"][a:e])
    assert np.array_equal(sub["best_hyp"].cpu().numpy(), out["best_hyp"][3:6])


def test_tracks_to_scales_without_poses(engine):
    """Correspondences alone -> raw scales (find_essential -> recover_pose -> fused stages 1-5), 10 % of the tracks replaced by
    mismatches: the scale lands within 3 % of the truth (the CPU chain of the same steps -- host build of the solver,
    oracle.extras.recover_pose, oracle.pipeline -- lands within 1 % on these frames) and close to the scale from the true poses."""
    import torch
    from mvoscalerecovery_b200 import synth
    b = synth.make_sequence(seed=21, n_frames=6, n_corr=1500, outlier_frac=0.1)
    rng = np.random.default_rng(2)
    ru, rv = b.ref_u.copy(), b.ref_v.copy()
    bad = rng.permutation(ru.size)[: ru.size // 10]
    ru[bad] = rng.uniform(0, 1241, bad.size).astype(np.float32); rv[bad] = rng.uniform(0, 376, bad.size).astype(np.float32)
    d = [_t(engine, x) for x in (b.offsets, b.cur_u, b.cur_v, ru, rv)]
    maxf = int(np.diff(b.offsets).max())
    out = engine.scale_frames_from_tracks(*d, max_features=maxf, hypotheses=128, threshold=0.5, seed=5)
    ref = engine.scale_frames_from_correspondences(*d[:3], _t(engine, b.ref_u), _t(engine, b.ref_v), _t(engine, b.poses), max_features=maxf, seed=5)
    torch.cuda.synchronize()
    raw, raw_true = out["raw_scale"].cpu().numpy(), ref["raw_scale"].cpu().numpy()
    n_inl = out["n_inliers"].cpu().numpy()
    truth_bad = np.zeros(ru.size, bool); truth_bad[bad] = True
    mask = out["e_mask"].cpu().numpy().astype(bool)
    assert (mask & truth_bad).sum() <= 0.02 * truth_bad.sum()                      # the mismatches are rejected
    assert (n_inl >= 0.85 * np.diff(b.offsets)).all()
    assert (out["status"].cpu().numpy() & 1).all()                                  # every frame produced a scale
    np.testing.assert_allclose(raw, b.true_scale, rtol=0.03)
    np.testing.assert_allclose(raw, raw_true, rtol=0.03)


def test_process_tracks_drop_in(engine, golden):
    """compat.vo_geometry.process_tracks = lines 129-147 of visual_odometry.py for one frame, numpy in / numpy out: the attributes
    the reference's main loop reads (motion_R, motion_t, feature3d, px_*_selected) are consistent with each other, with the
    batch entry points, and with the truth."""
    from mvoscalerecovery_b200.compat import vo_geometry
    z = golden
    Kmat = np.array([[K[0], 0, K[2]], [0, K[1], K[3]], [0, 0, 1.0]])
    off = z["offsets"]
    for f in (1, 6):
        a, e = off[f], off[f + 1]
        cur = np.stack([z["cur_u"][a:e], z["cur_v"][a:e]], 1); ref = np.stack([z["ref_u"][a:e], z["ref_v"][a:e]], 1)
        g = vo_geometry.process_tracks(cur, ref, Kmat, threshold=0.5, prob=0.999, max_iters=1000, seed=int(z["seed"]), frame=f, seq=int(z["seq"]))
        P = z["true_poses"][f].reshape(3, 4)
        tt = P[:, 3] / np.linalg.norm(P[:, 3])
        assert _angle_deg((np.trace(g["R"].T @ P[:, :3]) - 1) / 2) < 0.2 and _angle_deg(float(g["t"][:, 0] @ tt)) < 1.5
        assert abs(np.linalg.det(g["R"]) - 1) < 1e-9 and abs(np.linalg.norm(g["t"]) - 1) < 1e-9
        m = g["mask"]
        assert g["feature3d"].shape == (m.sum(), 3) and np.array_equal(g["px_cur_selected"], cur[m]) and np.array_equal(g["px_ref_selected"], ref[m])
        assert g["hyps_used"] == 128 and abs(g["n_inliers"] - int(z["n_inliers"][f])) <= 5
        truth = z["true_match"][a:e]
        assert (m & ~truth).sum() <= 4 and (m & truth).sum() >= 0.95 * truth.sum()
        X = g["feature3d"]
        assert (X[:, 2] > 0).all() and (X[:, 2] < 100).all()
        # the triangulated points reproject onto the current-image tracks (fx on both axes, main.py:102-104, fx == fy here)
        uv = np.stack([X[:, 0] * K[0] / X[:, 2] + K[2], X[:, 1] * K[1] / X[:, 2] + K[3]], 1)
        assert np.median(np.abs(uv - cur[m])) < 0.5
    with pytest.raises(RuntimeError):
        vo_geometry.process_tracks(cur[:4], ref[:4], Kmat)
