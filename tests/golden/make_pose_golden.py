"""Writes tests/golden/pose.npz: outputs of OpenCV's own cv2.findEssentialMat / cv2.recoverPose -- the calls of the
reference's VO step (src/thirdparty/MonocularVO/visual_odometry.py:129-133) -- on small seeded synthetic frames.
Run in the build container:  python tests/golden/make_pose_golden.py   (OpenCV 4.13 here; the algorithm lives in OpenCV,
not in the reference tree: calib3d five-point.cpp, recoverPose / decomposeEssentialMat)."""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mvoscalerecovery_b200 import synth          # noqa: E402


def skew(t):
    return np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])


def main():
    cam = synth.Camera()
    K = np.array([[cam.fx, 0, cam.cx], [0, cam.fy, cam.cy], [0, 0, 1.0]])
    b = synth.make_sequence(seed=77, n_frames=8, n_corr=400, outlier_frac=0.1)
    out = dict(offsets=b.offsets, cur_u=b.cur_u, cur_v=b.cur_v, ref_u=b.ref_u, ref_v=b.ref_v, true_poses=b.poses)
    Es, Rs, ts, counts, kinds = [], [], [], [], []
    for f in range(b.n_frames):
        a, e = b.offsets[f], b.offsets[f + 1]
        cur = np.stack([b.cur_u[a:e], b.cur_v[a:e]], 1).astype(np.float32)
        ref = np.stack([b.ref_u[a:e], b.ref_v[a:e]], 1).astype(np.float32)
        P = b.poses[f].reshape(3, 4)
        if f % 2 == 0:      # the essential matrix OpenCV's 5-point RANSAC finds, as the reference computes it
            E, _ = cv2.findEssentialMat(cur, ref, cameraMatrix=K, method=cv2.RANSAC, prob=0.999, threshold=0.5)
            E = E[:3]
            kinds.append(0)
        else:               # the exact essential matrix of the synthetic pose, arbitrary scale and sign
            t = P[:, 3] / np.linalg.norm(P[:, 3])
            E = skew(t) @ P[:, :3] * (-2.5 if f % 4 == 1 else 0.7)
            kinds.append(1)
        n, R, t, mask, _ = cv2.recoverPose(E, cur, ref, cameraMatrix=K, distanceThresh=100)
        Es.append(E.reshape(-1)); Rs.append(R.reshape(-1)); ts.append(t.reshape(-1)); counts.append(int((mask > 0).sum()))
    out.update(E=np.array(Es), R=np.array(Rs), t=np.array(ts), n_good=np.array(counts), kind=np.array(kinds))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "pose.npz"), **out)
    print("wrote pose.npz", out["n_good"], np.diff(b.offsets))
    for f in range(b.n_frames):
        P = b.poses[f].reshape(3, 4)
        print(f, "rot err", np.linalg.norm(out["R"][f].reshape(3, 3) - P[:, :3]), "t err", np.linalg.norm(out["t"][f] - P[:, 3] / np.linalg.norm(P[:, 3])))


if __name__ == "__main__":
    main()
