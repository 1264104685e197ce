"""Trajectory and scale evaluation: vectorised counterparts of the reference's post-hoc scripts
``script/evaluate_vo.py`` (KITTI segment errors) and ``script/evaluate_scale.py`` (per-frame scale error statistics).
Host-side numpy: a few thousand 4x4 products per sequence, run once after the GPU path has produced
``scales`` and ``poses`` (mvoscalerecovery_b200.offline).

    python -m mvoscalerecovery_b200.evaluate vo    ground_truth_poses.txt result_poses.txt
    python -m mvoscalerecovery_b200.evaluate scale ground_truth_scales.txt result_scales.txt
"""
from __future__ import annotations

import sys

import numpy as np

LENGTHS = (100, 200, 300, 400, 500, 600, 700, 800)        # segment lengths in metres (evaluate_vo.py:43)
STEP = 10                                                  # every 10th frame starts a segment (evaluate_vo.py:42)


def _mats(poses):
    p = np.asarray(poses, dtype=np.float64).reshape(-1, 3, 4)
    m = np.zeros((p.shape[0], 4, 4))
    m[:, :3] = p
    m[:, 3, 3] = 1.0
    return m


def trajectory_distances(poses):
    """Cumulative path length along the (N,12) poses (evaluate_vo.py:5-13)."""
    t = np.asarray(poses, dtype=np.float64)[:, 3:12:4]
    return np.concatenate([[0.0], np.cumsum(np.linalg.norm(t[:-1] - t[1:], axis=1))])


def sequence_errors(poses_gt, poses_result):
    """Rows [first_frame, rotation error / length (rad/m), translation error / length, length, speed] for every start frame
    (step 10) and every segment length that fits (evaluate_vo.py:37-76)."""
    gt, res = _mats(poses_gt), _mats(poses_result)
    dist = trajectory_distances(poses_gt)
    n = gt.shape[0]
    first = np.repeat(np.arange(0, n, STEP), len(LENGTHS))
    length = np.tile(np.asarray(LENGTHS, dtype=np.float64), (n + STEP - 1) // STEP)
    last = np.searchsorted(dist, dist[first] + length, side="right")       # first frame farther than `length` along the path
    ok = last < n
    first, last, length = first[ok], last[ok], length[ok]
    d_gt = np.linalg.inv(gt[first]) @ gt[last]
    d_res = np.linalg.inv(res[first]) @ res[last]
    err = np.linalg.inv(d_res) @ d_gt
    r_err = np.arccos(np.clip(0.5 * (err[:, 0, 0] + err[:, 1, 1] + err[:, 2, 2] - 1.0), -1.0, 1.0))
    t_err = np.linalg.norm(err[:, :3, 3], axis=1)
    speed = length / (0.1 * (last - first + 1))
    return np.stack([first.astype(np.float64), r_err / length, t_err / length, length, speed], 1)


def average_errors(errors):
    """Mean rotation error (deg/m) and translation error per segment length (evaluate_vo.py:77-96).  Returns
    (rot (L,), tra (L,), list of the per-length translation errors); lengths without a segment are skipped."""
    errors = np.asarray(errors, dtype=np.float64).reshape(-1, 5)
    rot, tra, tra_all = [], [], []
    for length in LENGTHS:
        sel = np.abs(errors[:, 3] - length) < 1
        tra_all.append(list(errors[sel, 2]))
        if sel.any():
            rot.append(errors[sel, 1].sum() / sel.sum())
            tra.append(errors[sel, 2].sum() / sel.sum())
    return np.array(rot) * 180 / np.pi, tra, tra_all


def kitti_translation_error(poses_gt, poses_result):
    """The headline number of the reference's README (average translation error over all segment lengths, as a fraction)."""
    _, tra, _ = average_errors(sequence_errors(poses_gt, poses_result))
    return float(np.mean(tra)) if len(tra) else float("nan")


def patch(data, window=10, step=2):
    """|mean| of the signed error over sliding windows (evaluate_scale.py:20-24)."""
    data = np.asarray(data, dtype=np.float64)
    starts = np.arange(0, data.shape[0] - window, step)
    if starts.size == 0:
        return np.zeros(0)
    c = np.concatenate([[0.0], np.cumsum(data)])
    return np.abs(c[starts + window] - c[starts]) / window


def scale_errors(gt, re):
    """The statistics evaluate_scale prints (evaluate_scale.py:4-14): dict(mean, max, within_0.1 ... within_0.5, windowed)."""
    gt, re = np.asarray(gt, dtype=np.float64), np.asarray(re, dtype=np.float64)
    n = re.shape[0]
    signed = gt[:n] - re
    er = np.abs(signed)
    out = dict(mean=float(np.mean(er)), max=float(np.max(er)))
    for thr in (0.1, 0.2, 0.3, 0.5):
        out["within_%g" % thr] = float(1 - np.sum(er > thr) / n)
    out["windowed"] = [float(np.mean(patch(signed, w, 10))) for w in (10, 20, 50, 100, 200, 300, 400, 500, 600, 700, 800)]
    return out


def filter(data, window=10):                               # noqa: A001  (the reference's name, evaluate_scale.py:25-29)
    """Causal running median over the last ``window`` values -- "filter_10".  Host restatement for small inputs; the batch
    path computes it on the GPU (mvosr_filter_sequences, filter10_out)."""
    data = np.asarray(data, dtype=np.float64)
    return np.array([data[0]] + [np.median(data[max(i - window + 1, 0):i + 1]) for i in range(1, data.shape[0])])


def main(argv=None):
    argv = sys.argv if argv is None else argv
    if len(argv) != 4 or argv[1] not in ("vo", "scale"):
        sys.exit("usage: python -m mvoscalerecovery_b200.evaluate vo|scale ground_truth.txt result.txt")
    gt, re = np.loadtxt(argv[2]), np.loadtxt(argv[3])
    if argv[1] == "vo":
        rot, tra, _ = average_errors(sequence_errors(gt, re))
        print(tra, rot)
        print(np.mean(tra), np.mean(rot))
    else:
        s = scale_errors(gt, re)
        print(s["mean"], s["max"], s["within_0.1"], s["within_0.2"], s["within_0.3"], s["within_0.5"])
        print(s["windowed"])


if __name__ == "__main__":
    main()
