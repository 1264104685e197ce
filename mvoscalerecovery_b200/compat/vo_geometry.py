"""The geometry of VisualOdometry.processFrame (src/thirdparty/MonocularVO/visual_odometry.py:129-147) on the GPU, numpy in /
numpy out: cv2.findEssentialMat + cv2.recoverPose + the dehomogenised, masked triangulation, i.e. what the reference's main loop
reads from the VO object afterwards (vo.motion_R, vo.motion_t, vo.feature3d, vo.px_cur_selected, vo.px_ref_selected;
src/main.py:87-104).  Three C-ABI calls (mvosr_find_essential_frames, mvosr_recover_pose_frames, mvosr_triangulate_frames) plus
mvosr_pose_mask_frames for the selection mask; no OpenCV and no CPU fallback.  A maintainer replaces lines 129-147 by

    g = vo_geometry.process_tracks(self.px_cur, self.px_ref, self.camera_matrix, frame=frame_id)
    self.motion_R, self.motion_t, self.feature3d = g["R"], g["t"], g["feature3d"]
    self.px_cur_selected, self.px_ref_selected = g["px_cur_selected"], g["px_ref_selected"]

Differences from OpenCV, by construction: the RANSAC draws from the Philox stream (seed, frame), not from cv::RNG, and checks
the adaptive count every 128 hypotheses; feature3d is float32-rounded (stage 1's output type)."""
import numpy as np

_ENGINES = {}


def _engine(K):
    from mvoscalerecovery_b200.batch import ScaleRecovery
    key = (float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2]))
    if key not in _ENGINES:
        _ENGINES[key] = ScaleRecovery(fx=key[0], fy=key[1], cx=key[2], cy=key[3])
    return _ENGINES[key]


def process_tracks(px_cur, px_ref, camera_matrix, threshold=0.5, prob=0.999, max_iters=1000, seed=0, frame=0, seq=0):
    """px_cur, px_ref: (n,2) pixel coordinates of the tracked features in the current / previous image.  Returns a dict:
    R (3,3), t (3,1) with x_ref = R x_cur + t, |t| = 1 (cv2.recoverPose's convention); E (3,3); mask (n,) bool = recoverPose's mask
    AND findEssentialMat's mask; feature3d (n',3) float64 in the current camera frame; px_cur_selected, px_ref_selected (n',2)."""
    import torch
    K = np.asarray(camera_matrix, dtype=np.float64).reshape(3, 3)
    eng = _engine(K)
    cur = np.ascontiguousarray(px_cur, dtype=np.float32).reshape(-1, 2)
    ref = np.ascontiguousarray(px_ref, dtype=np.float32).reshape(-1, 2)
    n = cur.shape[0]
    dev = eng.device
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev)
    off = t(np.array([0, n]), np.int32)
    fi = t(np.array([frame]), np.int32)
    d = [t(cur[:, 0], np.float32), t(cur[:, 1], np.float32), t(ref[:, 0], np.float32), t(ref[:, 1], np.float32)]
    ess = eng.find_essential_frames(off, *d, hypotheses=max_iters, threshold=threshold, seed=seed, frame_index=fi, seq_id=seq, confidence=prob)
    if int(ess["best_hyp"].cpu()[0]) < 0:
        raise RuntimeError("no essential matrix: fewer than five correspondences or a degenerate track set")   # cv2 returns an empty E here
    pose = eng.recover_pose_frames(off, *d, ess["essential"])                                   # the reference passes no mask to recoverPose
    mask = eng.pose_mask_frames(off, *d, pose["poses"], e_mask=ess["e_mask"])
    tri = eng.triangulate_frames(off, *d, pose["poses"], e_mask=ess["e_mask"])
    torch.cuda.synchronize(dev)
    m = int(tri["n_out"].cpu()[0])
    P = pose["poses"].cpu().numpy().reshape(3, 4)
    sel = mask.cpu().numpy().astype(bool)
    f3 = np.stack([tri[k][:m].cpu().numpy() for k in "xyz"], 1).astype(np.float64)
    assert int(sel.sum()) == m
    return dict(R=P[:, :3].copy(), t=P[:, 3:4].copy(), E=ess["essential"].cpu().numpy().reshape(3, 3), mask=sel, feature3d=f3,
                px_cur_selected=cur[sel], px_ref_selected=ref[sel], n_inliers=int(ess["n_inliers"].cpu()[0]),
                hyps_used=int(ess["hyps_used"].cpu()[0]))
