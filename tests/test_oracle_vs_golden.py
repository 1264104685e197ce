"""The oracle (oracle/pipeline.py, oracle/philox.py, oracle/exact_dt.c) pinned against outputs of the REFERENCE ITSELF
(tests/golden/*.npz, written by tests/golden/make_golden.py through oracle/ref_harness.py) and the known answers of
SURVEY.md section 8(c).  CPU only."""
import os

import numpy as np
import pytest

from oracle import pipeline as P
from oracle import philox

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_philox_known_answers():
    """Philox4x32-10 KATs (Random123 kat_vectors), SURVEY.md 8(c)."""
    assert philox.philox4x32_10((0, 0, 0, 0), (0, 0)) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    f = 0xffffffff
    assert philox.philox4x32_10((f, f, f, f), (f, f)) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert philox.philox4x32_10((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)


def test_philox_scalar_and_vector_streams_agree():
    seed = 0x1234567890abcdef
    for n in (3, 4, 17, 6726):
        hyps = np.arange(64)
        v = philox.sample3_positions_np(seed, hyps, 5, 2, n)
        for h in (0, 1, 63):
            assert tuple(int(x) for x in v[h]) == philox.sample3_positions(seed, h, 5, 2, n)
        assert np.all(v >= 0) and np.all(v < n)
        assert np.all(v[:, 0] != v[:, 1]) and np.all(v[:, 0] != v[:, 2]) and np.all(v[:, 1] != v[:, 2])


def test_graph_potential_known_answer():
    """graph.py __main__ (reference src/graph.py:156-165): column of v=[0,1,2], d=[2,1,1]; and the p=0.8 LUT."""
    tp = P.triangle_potential()
    assert tp[0].tolist() == [27, 9, 9, 3, 9, 3, 3, 1]
    assert tp[7].tolist() == [0, 0, 0, 0, 0, 0, 0, 64]
    v, d = np.array([0., 1., 2.]), np.array([2., 1., 1.])
    a = (v[0] - v[1]) * (d[0] - d[1]) < 0; b = (v[1] - v[2]) * (d[1] - d[2]) < 0; c = (v[0] - v[2]) * (d[0] - d[2]) < 0
    idx = 4 * int(a) + 2 * int(b) + int(c)
    assert tp[:, idx].tolist() == [3, 4, 4, 0, 12, 16, 16, 0]
    prob = P.vertex_probability_table()
    np.testing.assert_allclose(prob[idx], [0.8, 0.36363636, 0.36363636], rtol=1e-7)
    # p = 0.8 iff both edges at that vertex are consistent, else <= 0.528
    for i in range(8):
        a, b, c = (i >> 2) & 1, (i >> 1) & 1, i & 1
        want = [a and c, a and b, b and c]
        for k in range(3):
            assert (prob[i, k] > 0.6) == bool(want[k])
            assert abs(prob[i, k] - 0.8) < 1e-12 if want[k] else prob[i, k] <= 28.0 / 53.0 + 1e-12


def test_qhull_unit_square_known_answer():
    from scipy.spatial import Delaunay
    s = Delaunay(np.array([[0, 0], [1, 0], [1, 1], [0, 1]], float)).simplices
    assert P.canonicalise(s).tolist() == [[0, 1, 3], [1, 2, 3]]


def _frames(g):
    return [f for f in range(g.n_frames) if g.called(f)]


def test_oracle_intermediates_equal_reference(golden):
    """Every intermediate of rescale.ScaleEstimator.scale_calculation the reference produced, frame by frame."""
    g = golden
    n = 0
    for f in _frames(g):
        f3 = g.f3(f).astype(np.float64); f2 = g.f2(f).astype(np.float64)
        rec = P.frame_raw_scale(f3, f2, g.seed, f, 0, absolute_reference=1.7)
        sc = g.scalars(f)
        assert np.array_equal(rec["tri1"], g.get(f, "tri1")), "frame %d DT#1" % f
        assert np.array_equal(rec["keep"], g.get(f, "keep")), "frame %d keep" % f
        assert np.array_equal(rec["tri2"], g.get(f, "tri2")), "frame %d DT#2" % f
        flags = rec["loose"].astype(np.uint8) | (rec["tight"].astype(np.uint8) << 1) | (rec["valid"].astype(np.uint8) << 2)
        assert np.array_equal(flags, g.get(f, "flags")), "frame %d gates" % f
        np.testing.assert_allclose(rec["heights"], g.get(f, "heights"), rtol=1e-9)
        np.testing.assert_allclose(rec["height_level"], sc["height_level"], rtol=1e-12)
        assert np.array_equal(rec["data_id"], g.get(f, "data_id")), "frame %d vertex list" % f
        assert rec["updated"] == sc["updated"]
        if sc["updated"]:
            hyp = g.get(f, "hyp_log")
            assert rec["hyps_used"] == hyp.shape[0]
            assert rec["ic"] == sc["best_ic"]
            np.testing.assert_allclose(rec["height"], sc["height"], rtol=1e-9)
            np.testing.assert_allclose(rec["raw_scale"], sc["raw_scale"], rtol=1e-9)
            # the returned model's inlier set over the vertex list == the reference's own is_inlier (estimate_road_norm.py:17-18)
            ref_inl = np.unpackbits(g.get(f, "inlier"))[: rec["n_sel"]].astype(bool)
            aug = np.hstack([rec["point_selected"], np.ones((rec["n_sel"], 1))])
            assert np.array_equal(np.abs(aug @ rec["model"]) < 0.005, ref_inl), "frame %d inlier set" % f
            assert int(ref_inl.sum()) == sc["best_ic"]
        n += 1
    assert n > 0


def test_oracle_offline_loop_and_filter10_equal_reference(golden):
    """Driver gating (main_offline.py:57-88), slew limiter + deque median (rescale.py:168-178), filter_10."""
    g = golden
    f3s = [g.f3(f).astype(np.float64) for f in range(g.n_frames)]
    f2s = [g.f2(f).astype(np.float64) for f in range(g.n_frames)]
    scales, _ = P.offline_loop(f3s, f2s, g.z["move_flags"], g.seed, absolute_reference=1.7, window_size=5)
    np.testing.assert_allclose(scales, g.z["scales"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(P.filter10(g.z["scales"]), g.z["filter10"], rtol=0, atol=0)


def test_oracle_stage1_equals_opencv_golden(golden):
    """triangulate_dlt vs cv2.recoverPose's triangulation stored by the golden generator (float32 outputs)."""
    g = golden
    z = g.z
    from mvoscalerecovery_b200 import synth
    cam = synth.Camera()
    off = z["offsets"]
    checked = 0
    for f in range(g.n_frames):
        a, e = off[f], off[f + 1]
        cur = np.stack([z["cur_u"][a:e], z["cur_v"][a:e]], 1); ref = np.stack([z["ref_u"][a:e], z["ref_v"][a:e]], 1)
        Pm = z["poses"][f].reshape(3, 4)
        X, m = P.triangulate_dlt(cur, ref, Pm[:, :3], Pm[:, 3], cam.fx, cam.fy, cam.cx, cam.cy)
        f3 = g.f3(f)
        assert int(m.sum()) == f3.shape[0]
        got = X[m].astype(np.float32)
        ulp = np.abs(got.astype(np.float64) - f3.astype(np.float64)) / np.maximum(np.spacing(np.abs(f3)).astype(np.float64), 1e-30)
        assert ulp.size == 0 or ulp.max() <= 1.0
        checked += 1
    assert checked


def test_exact_delaunay_oracle_equals_qhull_on_generic_points():
    """oracle/exact_dt.c (exact arithmetic, symbolic tie-break) == canonicalised Qhull on points in general position."""
    from oracle import exact
    rng = np.random.default_rng(7)
    for n in (3, 10, 200, 1500):
        p = np.stack([rng.uniform(0, 1241, n), rng.uniform(186, 376, n)], 1).astype(np.float32)
        tri, dup = exact.delaunay_exact(p)
        assert not dup.any()
        assert np.array_equal(tri, P.delaunay_canonical(p.astype(np.float64)))
        ok, msg, _ = exact.validate_delaunay(p, tri, dup)
        assert ok, msg


def test_exact_delaunay_oracle_degenerate_inputs_are_valid():
    from oracle import exact
    grid = np.stack(np.meshgrid(np.arange(12.), 190 + np.arange(9.)), -1).reshape(-1, 2).astype(np.float32)
    tri, dup = exact.delaunay_exact(grid)
    assert tri.shape[0] == 2 * 11 * 8
    ok, msg, _ = exact.validate_delaunay(grid, tri, dup)
    assert ok, msg
    pts = np.array([[0, 200], [1, 200], [2, 200], [1, 200], [1.5, 201]], np.float32)
    tri, dup = exact.delaunay_exact(pts)
    assert dup.tolist() == [False, False, False, True, False]
    assert tri.tolist() == [[0, 1, 4], [1, 2, 4]]
    tri, dup = exact.delaunay_exact(np.array([[0, 200], [1, 200], [2, 200]], np.float32))
    assert tri.shape[0] == 0


def test_offline_loop_on_the_reference_mains_own_hand_off():
    """BASELINE configs[0] in small: src/main.py (unmodified) ran on rendered frames and wrote its hand-off file; the oracle's
    offline loop over that hand-off == what the reference's estimator makes of it (tests/golden/make_main_golden.py)."""
    from mvoscalerecovery_b200 import container
    z = np.load(os.path.join(ROOT, "tests", "golden", "main_c1.npz"))
    lists = container.unpack_sequence({k: z[k] for k in ("offsets", "move_flags", "motions", "x", "y", "z", "u", "v")})
    out = P.offline_loop(lists["feature3ds"], lists["feature2ds"], lists["move_flags"], int(z["seed"]),
                         absolute_reference=float(z["absolute_reference"]), window_size=5)
    scales = out[0] if isinstance(out, tuple) else out["scales"]
    np.testing.assert_allclose(scales, z["scales"], rtol=1e-9, atol=1e-12)
    assert z["called"].sum() >= 30 and np.median(np.abs(z["scales"][5:] - z["true_steps"][1:][5:len(z["scales"])]) / z["true_steps"][1:][5:len(z["scales"])]) < 0.1
