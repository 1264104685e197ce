"""The synthetic KITTI-shaped generator (the bench's data source): determinism, frame-range slicing (what the fleet
workload of bench.py builds per rank) and geometric sanity against the oracle's stage 1."""
import numpy as np

from mvoscalerecovery_b200 import synth


def test_frame_range_is_a_slice_of_the_full_sequence():
    a = synth.make_sequence(seed=3, n_frames=20, n_corr=300, seq=2, still_every=7)
    b = synth.make_sequence(seed=3, n_frames=20, n_corr=300, seq=2, still_every=7, frame_range=(5, 16))
    o = a.offsets
    for k in ("cur_u", "cur_v", "ref_u", "ref_v"):
        assert np.array_equal(getattr(b, k), getattr(a, k)[o[5]:o[16]])
    assert np.array_equal(b.poses, a.poses[5:16]) and np.array_equal(b.move_flags, a.move_flags[5:16])
    assert np.array_equal(np.diff(b.offsets), np.diff(o[5:17])) and np.array_equal(b.true_scale, a.true_scale[5:16])
    assert b.n_frames == 11 and b.offsets[0] == 0
    c = synth.make_sequence(seed=3, n_frames=20, n_corr=300, seq=2, still_every=7)
    assert np.array_equal(a.cur_u, c.cur_u)                                   # deterministic
    assert not np.array_equal(a.cur_u, synth.make_sequence(seed=3, n_frames=20, n_corr=300, seq=3, still_every=7).cur_u)


def test_geometry_recovers_the_camera_height():
    """Triangulating the synthetic correspondences with the oracle's DLT puts the inlier road points 1.7 m / scale below the camera."""
    from oracle import pipeline as P
    cam = synth.Camera()
    b = synth.make_sequence(seed=11, n_frames=3, n_corr=800, outlier_frac=0.0, pixel_noise=0.0)
    for f in range(3):
        a, e = b.offsets[f], b.offsets[f + 1]
        cur = np.stack([b.cur_u[a:e], b.cur_v[a:e]], 1); ref = np.stack([b.ref_u[a:e], b.ref_v[a:e]], 1)
        Pm = b.poses[f].reshape(3, 4)
        X, m = P.triangulate_dlt(cur, ref, Pm[:, :3], Pm[:, 3], cam.fx, cam.fy, cam.cx, cam.cy)
        uv = P.reproject(X[m], cam.fx, cam.cx, cam.cy)
        road = uv[:, 1] > 200
        rec = P.frame_raw_scale(X[m].astype(np.float32).astype(np.float64), uv.astype(np.float32).astype(np.float64), 1, f, 0, absolute_reference=1.7)
        assert road.sum() > 300 and abs(rec["raw_scale"] - b.true_scale[f]) < 0.01 * b.true_scale[f]
