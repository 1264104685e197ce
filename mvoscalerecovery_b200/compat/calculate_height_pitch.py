"""Importable Python-3 counterpart of the reference's batch script src/calculate_height_pitch.py (a Python-2 top-level
script: ``print`` statements, runs on import, reads sys.argv), with the same command line and the same output files:

    python calculate_height_pitch.py image_name_list_file feature_pos_path motion_path pose_path

reads ``<feature_pos_path><k>.txt`` (rows ``u v depth``) for k = 1, 2, ... and the N x 12 motion file, and writes
``result_heights_line_ransac.txt, refined_camera_height_means.txt, refined_camera_height_stds.txt,
refined_camera_height_t_means.txt, refined_pitch.txt, inlier_numbers.txt`` into the working directory
(calculate_height_pitch.py:19-32,228-244).

Per frame (calculate_height_pitch.py:62-204): back-project (u, v, depth), Delaunay over the pixels, per-triangle plane
n = P^-1 1 flipped to n_y >= 0 with signed height +-1/|n|, keep triangles whose pitch lies within (-95, -85) degrees of the
pitch predicted from the accumulated translations and whose height is positive, 500-hypothesis plane RANSAC at 0.005 over
their vertex list, inliers of all points at 0.01, then the "refinement": the plane through the FIRST THREE inliers
(``estimate`` uses xyzs[:3]), mean / std of the inliers' distances to it, and the mean height under the predicted pitch.

Here the frames are processed as one batch on the GPU through libmvosr.so: one mvosr_delaunay_frames launch for all
frames, one mvosr_triangle_planes launch for all triangles, one mvosr_ransac_planes launch for all vertex lists (Philox
stream: seed ``ransac_seed``, frame counter = the file number k); the cheap per-frame bookkeeping stays on the host.  The
image files named by the list are not read (the reference only draws on them).  Nothing is printed.
"""
import math
import os
import sys

import numpy as np

camera_focus = 718.856
camera_cx = 607.1928
camera_cy = 185.2157
PI_SCRIPT = 3.1415926                      # the script's own constant for rad -> deg (calculate_height_pitch.py:61,92)
ransac_seed = 0
RANSAC_ITERATIONS, RANSAC_THRESHOLD, INLIER_THRESHOLD, MIN_SELECTED = 500, 0.005, 0.01, 12
OUTPUT_FILES = ("result_heights_line_ransac.txt", "refined_camera_height_means.txt", "refined_camera_height_stds.txt",
                "refined_camera_height_t_means.txt", "refined_pitch.txt", "inlier_numbers.txt")


def back_project(points3d, cy=camera_cy):
    """(u, v, depth) rows -> camera coordinates (calculate_height_pitch.py:65-67)."""
    p = np.array(points3d, dtype=np.float64).reshape(-1, 3)
    p[:, 0] = p[:, 2] * (p[:, 0] - camera_cx) / camera_focus
    p[:, 1] = p[:, 2] * (p[:, 1] - cy) / camera_focus
    return p


def estimated_pitches(camera_motion_ts, n_frames):
    """get_pitch over translations 0..k for k = 1..n_frames (calculate_height_pitch.py:60)."""
    from estimate_road_norm import get_pitch
    ts = np.asarray(camera_motion_ts, dtype=np.float64)[:, 0:3]
    return np.array([get_pitch(ts[0:k + 1]) for k in range(1, n_frames + 1)])


def process_frames(frames, est_pitch, first_frame=1, seed=None):
    """All frames at once.  frames: list of (n,3) ``u v depth`` arrays (file k = first_frame + index); est_pitch: predicted
    pitch per frame in radians.  Returns a dict of per-frame arrays named after the output files' contents plus the
    intermediates the tests compare (``n_selected``, ``models``)."""
    import torch
    import _gpu
    from estimate_road_norm import estimate, get_inliers
    eng = _gpu.engine()
    dev = eng.device
    F = len(frames)
    seed = ransac_seed if seed is None else seed
    sizes = np.array([f.shape[0] for f in frames], np.int64)
    off = np.zeros(F + 1, np.int32)
    np.cumsum(sizes, out=off[1:])
    raw = np.concatenate([np.asarray(f, np.float64).reshape(-1, 3) for f in frames], 0) if F else np.zeros((0, 3))
    P3 = back_project(raw)
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev)
    # ---- Delaunay of every frame (one launch), triangles as global row indices
    dt = eng.delaunay_frames(t(off, np.int32), t(raw[:, 0], np.float32), t(raw[:, 1], np.float32), int(sizes.max()) if F else 1)
    n_tri = dt["n_tri"].cpu().numpy()
    tri_all = dt["tri"].cpu().numpy()
    tris = [tri_all[2 * off[f]: 2 * off[f] + n_tri[f]].astype(np.int64) for f in range(F)]
    gtri = np.concatenate([tr + off[f] for f, tr in enumerate(tris)], 0) if F else np.zeros((0, 3), np.int64)
    toff = np.zeros(F + 1, np.int64)
    np.cumsum(n_tri, out=toff[1:])
    # ---- planes of every triangle (one launch)
    pl = eng.triangle_planes(t(gtri, np.int32), t(P3, np.float64))
    normal = pl["normal"].cpu().numpy(); h = pl["height"].cpu().numpy()
    ln = np.sqrt(np.sum(normal * normal, 1))
    flip = normal[:, 1] < 0                                  # n_y < 0: flip the normal, the height goes negative (:84-86)
    height = np.where(flip, -h, h)
    pitch_deg = np.arcsin(-np.abs(normal[:, 1]) / ln) * 180 / PI_SCRIPT
    # ---- triangle gates and vertex lists (:101-131)
    lists, sel_off = [], np.zeros(F + 1, np.int32)
    for f in range(F):
        a, b = toff[f], toff[f + 1]
        e = est_pitch[f] * 180 / PI_SCRIPT
        ok = (pitch_deg[a:b] > e - 95) & (pitch_deg[a:b] < e - 85) & (height[a:b] > 0)
        ids = tris[f][ok].reshape(-1)
        lists.append(ids)
        sel_off[f + 1] = sel_off[f] + (ids.shape[0] if ids.shape[0] >= MIN_SELECTED else 0)
    sel_pts = np.concatenate([P3[off[f] + ids] for f, ids in enumerate(lists) if ids.shape[0] >= MIN_SELECTED] + [np.zeros((0, 3))], 0)
    # ---- plane RANSAC of every vertex list (one launch); Philox frame counter = the file number
    rs = eng.ransac_planes(t(sel_off, np.int32), t(sel_pts, np.float64), iterations=RANSAC_ITERATIONS, threshold=RANSAC_THRESHOLD,
                           seed=seed, frame_index=t(np.arange(first_frame, first_frame + F), np.int32))
    models = rs["model"].cpu().numpy()
    # ---- per-frame bookkeeping (:139-204); the previous frame's model and inliers carry over when a list is too short
    out = {k: np.zeros(F) for k in ("heights", "h_means", "h_stds", "h_t_means", "pitches", "inlier_numbers")}
    out["n_selected"] = np.array([ids.shape[0] for ids in lists])
    out["models"] = models
    norm_norm, inliers = 1.0, None
    for f in range(F):
        pts = P3[off[f]:off[f + 1]]
        if lists[f].shape[0] >= MIN_SELECTED:
            m = models[f]
            inliers = pts[get_inliers(m, pts, INLIER_THRESHOLD), :]
            n3, h_bar = m[:3], -m[3]
            if n3[1] < 0:
                n3, h_bar = -n3, -h_bar
            norm_norm = math.sqrt(float(n3 @ n3)) / h_bar
        if inliers is None:
            raise ValueError("frame %d: fewer than %d selected points and no earlier frame to fall back on" % (first_frame + f, MIN_SELECTED))
        out["heights"][f] = 1 / norm_norm
        out["inlier_numbers"][f] = inliers.shape[0]
        r = np.array(estimate(inliers))[:3]                  # plane through the first three inliers (:177-186)
        if r[1] < 0:
            r = -r
        r = r / math.sqrt(float(r @ r))
        out["pitches"][f] = math.asin(r[1])
        hs = inliers @ r
        out["h_means"][f], out["h_stds"][f] = np.mean(hs), np.std(hs)
        out["h_t_means"][f] = np.mean(inliers[:, 2] * math.sin(est_pitch[f]) + inliers[:, 1] * math.cos(est_pitch[f]))
    return out


def load_frames(feature_pos_path, limit=4540):
    """``<feature_pos_path><k>.txt`` for k = 1, 2, ... up to the first missing file (the reference hard-codes 4540 frames)."""
    frames = []
    while len(frames) < limit:
        name = feature_pos_path + str(len(frames) + 1) + ".txt"
        if not os.path.isfile(name):
            break
        frames.append(np.loadtxt(name).reshape(-1, 3))
    return frames


def main(argv=None):
    argv = sys.argv if argv is None else argv
    if len(argv) < 5:
        sys.exit("python calculate_height_pitch.py image_name_list_file feature_pos_path motion_path pose_path")
    feature_pos_path, motion_path = argv[2], argv[3]
    camera_motion_ts = np.loadtxt(motion_path)[:, 3::4]
    frames = load_frames(feature_pos_path, limit=min(4540, camera_motion_ts.shape[0] - 1))
    res = process_frames(frames, estimated_pitches(camera_motion_ts, len(frames)))
    for name, key in zip(OUTPUT_FILES, ("heights", "h_means", "h_stds", "h_t_means", "pitches", "inlier_numbers")):
        np.savetxt(name, res[key])
    return res


if __name__ == "__main__":
    main()
