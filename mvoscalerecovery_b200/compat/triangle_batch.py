"""Importable Python-3 counterpart of the reference's batch script src/triangle_batch.py (Python 2, runs on import), same
command line:

    python triangle_batch.py image_name_list_file feature_pos_path

For every ``<feature_pos_path><k>.txt`` (rows ``u v depth``), k = 1, 2, ...: Delaunay over the pixels, per-triangle plane
n = P^-1 1 of the back-projected vertices, keep the triangles with n_y/|n| > 0.98 and positive mean height, clip at three
standard deviations and print the mean height (triangle_batch.py:23-63).  Note the script's own principal point:
cy = 182.2157 (triangle_batch.py:22), not 185.2157.

All frames are processed as one batch on the GPU through libmvosr.so (mvosr_delaunay_frames + mvosr_triangle_planes); the
per-frame statistics are a handful of numpy reductions.  The reference also saves one matplotlib figure per frame
(``result/result<k>.png``): GUI output, not provided.
"""
import os
import sys

import numpy as np

camera_focus = 718.856
camera_cx = 607.1928
camera_cy = 182.2157
PI_SCRIPT = 3.1415926


def triangle_data(frames):
    """Per frame the (T,3) rows [n_y/|n|, |pitch| in degrees, mean height] of triangle_batch.py:31-44, canonical triangle order."""
    import torch
    import _gpu
    eng = _gpu.engine()
    dev = eng.device
    F = len(frames)
    sizes = np.array([f.shape[0] for f in frames], np.int64)
    off = np.zeros(F + 1, np.int32)
    np.cumsum(sizes, out=off[1:])
    raw = np.concatenate([np.asarray(f, np.float64).reshape(-1, 3) for f in frames], 0) if F else np.zeros((0, 3))
    P3 = raw.copy()
    P3[:, 0] = raw[:, 2] * (raw[:, 0] - camera_cx) / camera_focus
    P3[:, 1] = raw[:, 2] * (raw[:, 1] - camera_cy) / camera_focus
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev)
    dt = eng.delaunay_frames(t(off, np.int32), t(raw[:, 0], np.float32), t(raw[:, 1], np.float32), int(sizes.max()) if F else 1)
    n_tri = dt["n_tri"].cpu().numpy()
    tri_all = dt["tri"].cpu().numpy()
    gtri = np.concatenate([tri_all[2 * off[f]: 2 * off[f] + n_tri[f]].astype(np.int64) + off[f] for f in range(F)], 0) if F else np.zeros((0, 3), np.int64)
    pl = eng.triangle_planes(t(gtri, np.int32), t(P3, np.float64))
    normal = pl["normal"].cpu().numpy(); mean_y = pl["mean_y"].cpu().numpy()
    s = normal[:, 1] / np.sqrt(np.sum(normal * normal, 1))
    data = np.stack([s, np.abs(np.arcsin(s)) * 180 / PI_SCRIPT, mean_y], 1)
    toff = np.concatenate([[0], np.cumsum(n_tri)])
    return [data[toff[f]:toff[f + 1]] for f in range(F)]


def clipped_mean_height(data):
    """Mean height of the road-like triangles after 3-sigma clipping (triangle_batch.py:53-62)."""
    d = data[data[:, 0] > 0.98]
    d = d[d[:, 2] > 0]
    mean, std = np.mean(d[:, 2]), np.std(d[:, 2])
    d = d[d[:, 2] > mean - 3 * std]
    d = d[d[:, 2] < mean + 3 * std]
    return np.mean(d[:, 2])


def frame_heights(frames):
    return np.array([clipped_mean_height(d) for d in triangle_data(frames)])


def main(argv=None):
    argv = sys.argv if argv is None else argv
    if len(argv) < 3:
        sys.exit("python triangle_batch.py image_name_list_file feature_pos_path")
    frames = []
    while len(frames) < 4540 and os.path.isfile(argv[2] + str(len(frames) + 1) + ".txt"):
        frames.append(np.loadtxt(argv[2] + str(len(frames) + 1) + ".txt").reshape(-1, 3))
    heights = frame_heights(frames)
    for h in heights:
        print(h)
    return heights


if __name__ == "__main__":
    main()
