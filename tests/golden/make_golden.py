"""Generate the golden vectors under tests/golden/ by RUNNING THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Every output below comes from the unmodified reference modules (src/rescale.py, src/graph.py,
src/estimate_road_norm.py, src/thirdparty/Ransac/ransac.py, script/evaluate_scale.py) or from the
third-party call the reference makes (cv2.recoverPose), executed through oracle/ref_harness.py
(deterministic shims only: Philox sampler, canonicalised Delaunay order, matplotlib stub).
The inputs are stored with the outputs (float32 bit patterns), so the tests never regenerate
them.  Seeds are fixed; re-running reproduces the files bit for bit.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from mvoscalerecovery_b200 import synth                      # noqa: E402
from oracle import ref_harness as H                          # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
SEED = 20261017            # Philox hypothesis-stream seed used for every golden
CAM = synth.Camera()


def reference_stage1(cur, ref, R, t):
    """cv2.recoverPose(E, px_cur, px_ref, K, distanceThresh=100) exactly as the reference calls it
    (src/thirdparty/MonocularVO/visual_odometry.py:132-147) with E built from the known pose."""
    import cv2
    K = np.eye(3)
    K[0, 0] = K[1, 1] = CAM.fx
    K[0, 2], K[1, 2] = CAM.cx, CAM.cy
    tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
    E = tx @ R
    _, R_est, t_est, mask, pts4 = cv2.recoverPose(E, cur, ref, cameraMatrix=K, distanceThresh=100)
    mask_bool = np.array(mask > 0).reshape(-1)
    X = (pts4[:3] / pts4[3:4]).T
    return R_est, t_est.reshape(-1), mask_bool, X


def frames_from_batch(b):
    """Stage 1 through the reference's OpenCV call; outputs rounded to float32 (the bit pattern both
    sides consume downstream), feature2d by the formula of src/main.py:102-104."""
    f3s, f2s, s1 = [], [], []
    for f in range(b.n_frames):
        a, e = b.offsets[f], b.offsets[f + 1]
        if not b.move_flags[f]:
            f3s.append(np.zeros((0, 3))); f2s.append(np.zeros((0, 2))); s1.append(None)
            continue
        cur = np.stack([b.cur_u[a:e], b.cur_v[a:e]], 1)
        ref = np.stack([b.ref_u[a:e], b.ref_v[a:e]], 1)
        P = b.poses[f].reshape(3, 4)
        R_est, t_est, mask, X = reference_stage1(cur, ref, P[:, :3], P[:, 3])
        assert np.allclose(R_est, P[:, :3], atol=1e-9) and np.allclose(t_est, P[:, 3], atol=1e-9), "recoverPose picked another pose"
        Xm = X[mask]
        uv = Xm[:, 0:2].copy()
        uv[:, 0] = uv[:, 0] * CAM.fx / Xm[:, 2] + CAM.cx
        uv[:, 1] = uv[:, 1] * CAM.fx / Xm[:, 2] + CAM.cy
        f3s.append(Xm.astype(np.float32).astype(np.float64))
        f2s.append(uv.astype(np.float32).astype(np.float64))
        s1.append(dict(mask=mask, X=X, R_est=R_est, t_est=t_est))
    return f3s, f2s, s1


def pack_sequence(name, b, store_stage1_frames=()):
    ns = H.load_reference(seed=SEED)
    f3s, f2s, s1 = frames_from_batch(b)
    scales, recs = H.run_offline_loop(ns, f3s, f2s, b.move_flags, absolute_reference=1.7, window_size=5, seq=0)
    out = dict(seed=np.uint64(SEED), n_frames=np.int32(b.n_frames), move_flags=b.move_flags,
               scales=scales, filter10=H.reference_filter10(scales), true_scale=b.true_scale,
               offsets=b.offsets, cur_u=b.cur_u, cur_v=b.cur_v, ref_u=b.ref_u, ref_v=b.ref_v, poses=b.poses)
    for f in range(b.n_frames):
        out["f%d_f3" % f] = f3s[f].astype(np.float32)
        out["f%d_f2" % f] = f2s[f].astype(np.float32)
        r = recs[f]
        out["f%d_called" % f] = np.bool_(r is not None)
        if f in store_stage1_frames and s1[f] is not None:
            out["f%d_s1_mask" % f] = s1[f]["mask"]
            out["f%d_s1_X" % f] = s1[f]["X"]
        if r is None:
            continue
        for k in ("tri1", "tri2"):
            out["f%d_%s" % (f, k)] = r[k].astype(np.int16 if r[k].max(initial=0) < 32768 else np.int32)
        out["f%d_keep" % f] = r["keep"]
        out["f%d_flags" % f] = (r["loose"].astype(np.uint8) | (r["tight"].astype(np.uint8) << 1) | (r["valid"].astype(np.uint8) << 2))
        out["f%d_heights" % f] = r["heights"]
        out["f%d_pitch_deg" % f] = r["pitch_deg"]
        out["f%d_data_id" % f] = r["data_id"].astype(np.int16 if r["data_id"].size == 0 or r["data_id"].max() < 32768 else np.int32)
        out["f%d_hyp_log" % f] = r["hyp_log"].astype(np.int32)
        out["f%d_model" % f] = r["model"]
        out["f%d_inlier" % f] = np.packbits(r["inlier"])          # is_inlier(model, data[j]) over the vertex list (N_sel bits)
        out["f%d_scalars" % f] = np.array([r["height_level"], r["best_ic"], r["raw_scale"], r["height"], float(r["updated"]),
                                           r["state_before"], r["state_after"], r["scale_out"], float(r["second_dt"])])
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024), "scales[:5] =", scales[:5])


def pack_unrounded(name, b):
    """The reference on the float64 values its own front-end hands over (cv2.recoverPose's output and main.py:102-104's
    reprojection, NOT rounded to float32): what a caller of the per-frame drop-in actually passes.  The CUDA path stages
    float32; this golden quantifies what that rounding changes (tests/test_f64_handoff.py)."""
    ns = H.load_reference(seed=SEED)
    f3s, f2s = [], []
    for f in range(b.n_frames):
        a, e = b.offsets[f], b.offsets[f + 1]
        cur = np.stack([b.cur_u[a:e], b.cur_v[a:e]], 1)
        ref = np.stack([b.ref_u[a:e], b.ref_v[a:e]], 1)
        P = b.poses[f].reshape(3, 4)
        _, _, mask, X = reference_stage1(cur, ref, P[:, :3], P[:, 3])
        Xm = X[mask]
        uv = Xm[:, 0:2].copy()
        uv[:, 0] = uv[:, 0] * CAM.fx / Xm[:, 2] + CAM.cx
        uv[:, 1] = uv[:, 1] * CAM.fx / Xm[:, 2] + CAM.cy
        f3s.append(Xm); f2s.append(uv)
    scales, recs = H.run_offline_loop(ns, f3s, f2s, b.move_flags, absolute_reference=1.7, window_size=5, seq=0)
    out = dict(seed=np.uint64(SEED), n_frames=np.int32(b.n_frames), scales=scales)
    for f in range(b.n_frames):
        r = recs[f]
        out["f%d_f3" % f] = f3s[f]; out["f%d_f2" % f] = f2s[f]                    # float64
        out["f%d_scalars" % f] = np.array([r["raw_scale"], r["best_ic"], r["data_id"].shape[0], int(r["keep"].sum()), r["tri2"].shape[0],
                                           r["height_level"], r["scale_out"]])
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024), "scales[:5] =", scales[:5])


def main():
    # G0: the unrounded float64 hand-off (8 frames of the headline shape)
    pack_unrounded("seq_f64", synth.make_sequence(seed=404, n_frames=8, n_corr=2500, outlier_frac=0.10))
    # G1: the headline shape -- ~2.5k correspondences / ~2k ROI features per frame, 10 % outliers
    b = synth.make_sequence(seed=101, n_frames=10, n_corr=2500, outlier_frac=0.10)
    pack_sequence("seq_2k", b, store_stage1_frames=(0, 5))
    # G2: small frames, heavy outliers, a still frame and frames under the n>100 gate
    b = synth.make_sequence(seed=202, n_frames=48, n_corr=420, outlier_frac=0.30, n_jitter=0.6, still_every=11)
    pack_sequence("seq_small", b, store_stage1_frames=(1,))
    # G3: clean large steps (exercises the +-0.3 slew limiter from the initial state 1)
    sc = 1.9 + 0.45 * np.sin(np.arange(16) * 0.9)
    b = synth.make_sequence(seed=303, n_frames=16, n_corr=900, outlier_frac=0.0, pixel_noise=0.02, scales=sc)
    pack_sequence("seq_clean", b)


if __name__ == "__main__":
    main()
