"""CPU restatement of the reference's per-frame scale-recovery path -- TEST INFRASTRUCTURE.

This is the ORACLE: a numpy/scipy restatement (vectorised, no per-triangle Python loops)
of what ``/root/reference/src`` computes for the hot path of SURVEY.md section 8(a).
It exists to check the CUDA path; only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline legs may import it.  The product never routes through it.

Pinning: the reference holds no tests or golden vectors for this path (SURVEY.md
section 4), so the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, generated in
the build container by ``tests/golden/make_golden.py`` through ``oracle/ref_harness.py``
and committed under ``tests/golden/`` (checked by ``tests/test_oracle_vs_golden.py``),
plus the known answers of SURVEY.md section 8(c) (graph.py __main__, Philox KATs, Qhull unit
square).  Third-party arithmetic the reference calls (not under /root/reference, all
present in this image, no version pins in the reference): Qhull 8.0.2 via
scipy.spatial.Delaunay 1.18.1; LAPACK via numpy 2.3.5 (inv, svd); OpenCV 4.13.0
triangulatePoints/recoverPose.  The oracle calls the same libraries where the reference
does, so those steps are bit-identical rather than re-derived.

Every function cites the reference lines it follows.
"""
from __future__ import annotations

from collections import deque

import numpy as np
from scipy.spatial import Delaunay

from oracle.philox import sample3_positions_np

VANISH = 185                     # src/rescale.py:30
EDGE_POTENTIAL = [[3, 1], [2, 2], [2, 2], [0, 4]]     # src/rescale.py:32
RANSAC_ITERS = 100               # src/rescale.py:155
RANSAC_THR = 0.005               # src/rescale.py:155
MIN_SEL = 12                     # src/rescale.py:152
SLEW = 0.3                       # src/rescale.py:169-172
MIN_FEATURES = 100               # src/param.py:37 minimum_feature_for_scale


# ----------------------------------------------------------------------------- stage 1
def triangulate_dlt(cur_uv, ref_uv, R, t, fx, fy, cx, cy, dist=100.0):
    """Linear triangulation as done inside cv2.recoverPose(E, px_cur, px_ref, K, distanceThresh=100)
    (src/thirdparty/MonocularVO/visual_odometry.py:129-147): normalise by K in float64, build
    A (4x4) with rows x*P[2]-P[0], y*P[2]-P[1] for P0=[I|0] (current) and P1=[R|t] (ref), take the
    last right-singular vector, dehomogenise; mask = cheirality and depth<dist in both views.
    Returns (X (n,3) f64 in the current camera frame, mask (n,) bool)."""
    cur = np.asarray(cur_uv, dtype=np.float64)
    ref = np.asarray(ref_uv, dtype=np.float64)
    x0 = (cur[:, 0] - cx) / fx
    y0 = (cur[:, 1] - cy) / fy
    x1 = (ref[:, 0] - cx) / fx
    y1 = (ref[:, 1] - cy) / fy
    P0 = np.hstack([np.eye(3), np.zeros((3, 1))])
    P1 = np.hstack([np.asarray(R, dtype=np.float64).reshape(3, 3), np.asarray(t, dtype=np.float64).reshape(3, 1)])
    n = cur.shape[0]
    A = np.empty((n, 4, 4))
    A[:, 0, :] = x0[:, None] * P0[2] - P0[0]
    A[:, 1, :] = y0[:, None] * P0[2] - P0[1]
    A[:, 2, :] = x1[:, None] * P1[2] - P1[0]
    A[:, 3, :] = y1[:, None] * P1[2] - P1[1]
    _, _, Vt = np.linalg.svd(A)
    Q = Vt[:, 3, :]                                    # (n,4) homogeneous, unit 4-norm, sign arbitrary
    mask = (Q[:, 2] * Q[:, 3]) > 0
    with np.errstate(divide="ignore", invalid="ignore"):
        X = Q[:, :3] / Q[:, 3:4]
    mask &= X[:, 2] < dist
    z1 = X @ P1[2, :3] + P1[2, 3]
    mask &= (z1 > 0) & (z1 < dist)
    return X, mask


def reproject(X, fx, cx, cy):
    """src/main.py:102-104 -- note fx on BOTH axes."""
    u = X[:, 0] * fx / X[:, 2] + cx
    v = X[:, 1] * fx / X[:, 2] + cy
    return np.stack([u, v], axis=1)


# ----------------------------------------------------------------------------- stage 2
def canonicalise(simplices):
    s = np.sort(np.asarray(simplices, dtype=np.int64), axis=1)
    order = np.lexsort((s[:, 2], s[:, 1], s[:, 0]))
    return s[order].astype(np.int32)


def delaunay_canonical(points2d):
    """scipy.spatial.Delaunay (Qhull, the reference's own dependency, src/rescale.py:124,136)
    with the simplex list canonicalised (rows ascending, then lexsorted)."""
    return canonicalise(Delaunay(np.asarray(points2d, dtype=np.float64)).simplices)


def triangle_potential(edge_potential=EDGE_POTENTIAL):
    """src/graph.py:6-17 / :110-122."""
    ep = np.array(edge_potential, dtype=np.float64)
    tp = np.ones((8, 8))
    for row in range(8):
        r = [int(row & 4 != 0), int(row & 2 != 0), int(row & 1 != 0)]
        for col in range(8):
            c = [int(col & 4 != 0), int(col & 2 != 0), int(col & 1 != 0)]
            tp[row, col] = ep[r[0] * 2 + r[1], c[0]] * ep[r[1] * 2 + r[2], c[1]] * ep[r[0] * 2 + r[2], c[2]]
    return tp


def vertex_probability_table(edge_potential=EDGE_POTENTIAL):
    """(8,3) table p[idx, k] of src/graph.py:134-145 for every observation index."""
    tp = triangle_potential(edge_potential)
    rng = np.arange(8)
    sel = [(rng & 4) > 0, (rng & 2) > 0, (rng & 1) > 0]
    out = np.empty((8, 3))
    for idx in range(8):
        pot = tp[:, idx]
        z = np.sum(pot)
        for k in range(3):
            out[idx, k] = np.sum(pot[sel[k]]) / z
    return out


def graph_keep(tri, v, d, edge_potential=EDGE_POTENTIAL):
    """GraphChecker.find_inliers (src/graph.py:18-36): per triangle the three sign tests of
    check_triangle (:124-129, strict <0), per vertex keep = (#incident triangles with p>0.6) /
    (#incident triangles) > 0.5; a vertex in no triangle gives 0/0 = nan -> False."""
    tri = np.asarray(tri, dtype=np.int64)
    n = v.shape[0]
    vt = v[tri]
    dt = d[tri]
    a = ((vt[:, 0] - vt[:, 1]) * (dt[:, 0] - dt[:, 1]) < 0).astype(np.int64)
    b = ((vt[:, 1] - vt[:, 2]) * (dt[:, 1] - dt[:, 2]) < 0).astype(np.int64)
    c = ((vt[:, 0] - vt[:, 2]) * (dt[:, 0] - dt[:, 2]) < 0).astype(np.int64)
    idx = a * 4 + b * 2 + c
    ptab = vertex_probability_table(edge_potential)
    passed = ptab[idx] > 0.6                        # (T,3)
    total = np.bincount(tri.reshape(-1), minlength=n)
    good = np.bincount(tri.reshape(-1), weights=passed.reshape(-1).astype(np.float64), minlength=n)
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = good / total
    return ratio > 0.5


# ----------------------------------------------------------------------------- stage 3
def flat_selection(feature3d, tri):
    """src/rescale.py:75-102: n = P^-1 . 1 per triangle (LAPACK inverse, as np.matrix.I),
    pitch_deg = asin(-n_y/|n|)*180/pi, height = 1/|n|, loose = pitch<-80, tight = pitch<-85,
    height_level = 0.9*median(height[loose]), valid = tight & height>height_level.
    Returns dict with everything the harness records."""
    tri = np.asarray(tri, dtype=np.int64)
    b = np.ones((3, 1), float)
    P = feature3d[tri]                                # (T,3,3) rows = points
    Pi = np.linalg.inv(P)
    normals = (Pi @ b).reshape(-1, 3)
    nlen = np.sqrt(np.sum(normals * normals, 1)).reshape(-1, 1)
    normals = normals / nlen
    pitch_deg = np.arcsin(-normals[:, 1]) * 180 / np.pi
    loose = pitch_deg < -80
    tight = pitch_deg < -85
    heights = (1 / nlen).reshape(-1)
    with np.errstate(invalid="ignore"):
        height_level = 0.9 * (np.median(heights[loose]))
    valid = tight & (heights > height_level)
    data_id = tri[valid].reshape(-1)
    return dict(pitch_deg=pitch_deg, heights=heights, loose=loose, tight=tight, valid=valid,
                height_level=float(height_level), data_id=data_id.astype(np.int32),
                loose_heights=heights[loose])


def feature_selection(feature3d, feature2d, edge_potential=EDGE_POTENTIAL, vanish=VANISH):
    """src/rescale.py:113-148."""
    roi = feature2d[:, 1] > vanish
    f2 = feature2d[roi, :]
    f3 = feature3d[roi, :]
    tri1 = delaunay_canonical(f2)
    keep = graph_keep(tri1, f2[:, 1], f3[:, 2], edge_potential)
    second = bool(np.sum(keep) > 10)
    if second:
        f2 = f2[keep, :]
        f3 = f3[keep, :]
        tri2 = delaunay_canonical(f2)
    else:
        tri2 = tri1
    fs = flat_selection(f3, tri2)
    pts = f3[fs["data_id"]]
    out = dict(roi=roi, tri1=tri1, keep=keep, second_dt=second, tri2=tri2, f3_sel=f3, f2_sel=f2,
               point_selected=pts)
    out.update(fs)
    return out


# ----------------------------------------------------------------------------- stage 4
def estimate_plane_svd(p3):
    """estimate() of src/estimate_road_norm.py:13-15: last right-singular vector of [p 1] (3x4)."""
    a = np.ones((3, 4))
    a[:, :3] = p3
    return np.linalg.svd(a)[-1][-1, :]


def ransac_plane(points, seed, frame, seq=0, max_iterations=RANSAC_ITERS, thr=RANSAC_THR,
                 stop_at_goal=True):
    """get_pitch_ransac + run_ransac (src/estimate_road_norm.py:66-70,
    src/thirdparty/Ransac/ransac.py:3-23) with the Philox position stream of oracle/philox.py:
    sequential hypotheses, ic = #{|m.[x 1]| < thr} (strict), keep the first strictly larger ic,
    stop at the first ic > 0.8*N.  Returns dict(model, ic, best_hyp, hyps_used, positions,
    ics, degenerate)."""
    pts = np.asarray(points, dtype=np.float64)
    n = pts.shape[0]
    goal = n * 0.8
    pos = sample3_positions_np(seed, np.arange(max_iterations), frame, seq, n)
    aug = np.ones((n, 4))
    aug[:, :3] = pts
    best_ic, best_m, best_h = 0, None, -1
    ics = []
    degenerate = []
    used = 0
    for h in range(max_iterations):
        s = pts[pos[h]]
        degenerate.append(bool(np.linalg.matrix_rank(np.hstack([s, np.ones((3, 1))])) < 3))
        m = estimate_plane_svd(s)
        ic = int(np.sum(np.abs(aug @ m) < thr))
        ics.append(ic)
        used = h + 1
        if ic > best_ic:
            best_ic, best_m, best_h = ic, m, h
            if ic > goal and stop_at_goal:
                break
    return dict(model=best_m, ic=best_ic, best_hyp=best_h, hyps_used=used, positions=pos[:used],
                ics=np.asarray(ics), degenerate=np.asarray(degenerate))


def height_from_model(m):
    """src/rescale.py:156-166: n=m[:3], h_bar=-m[3]; flip both if n_y<0; height = h_bar/|n|."""
    n = np.array(m[:3], dtype=np.float64)
    h_bar = -float(m[3])
    if n[1] < 0:
        n = -n
        h_bar = -h_bar
    norm_norm = np.sqrt(float(n @ n)) / h_bar
    return 1.0 / norm_norm


# ----------------------------------------------------------------------------- stage 5/6
class TemporalState:
    """Slew limiter + deque median of src/rescale.py:23-35,168-178."""

    def __init__(self, window_size=6):
        self.scale = 1
        self.queue = deque()
        self.window_size = window_size

    def step(self, raw_scale, updated):
        if updated:
            if raw_scale - self.scale > SLEW:
                self.scale += SLEW
            elif raw_scale - self.scale < -SLEW:
                self.scale -= SLEW
            else:
                self.scale = raw_scale
        self.queue.append(self.scale)
        if len(self.queue) > self.window_size:
            self.queue.popleft()
        return float(np.median(self.queue))


def filter10(data, window=10):
    """script/evaluate_scale.py:25-29 ("filter_10"): causal running median."""
    data = np.asarray(data, dtype=np.float64)
    out = [data[0]]
    for i in range(1, data.shape[0]):
        out.append(np.median(data[max(i - window + 1, 0):i + 1]))
    return np.array(out)


def frame_raw_scale(feature3d, feature2d, seed, frame, seq=0, absolute_reference=1.7,
                    max_iterations=RANSAC_ITERS, thr=RANSAC_THR, edge_potential=EDGE_POTENTIAL):
    """Stages 2..5 for one frame: everything of ScaleEstimator.scale_calculation
    (src/rescale.py:191-193) except the temporal state. Returns a record dict."""
    rec = feature_selection(feature3d, feature2d, edge_potential)
    pts = rec["point_selected"]
    rec["n_sel"] = int(pts.shape[0])
    rec["updated"] = bool(pts.shape[0] >= MIN_SEL)
    if rec["updated"]:
        rr = ransac_plane(pts, seed, frame, seq, max_iterations, thr)
        rec.update(rr)
        rec["height"] = height_from_model(rr["model"])
        rec["raw_scale"] = absolute_reference / rec["height"]
    else:
        rec["raw_scale"] = np.nan
        rec["height"] = np.nan
    return rec


def offline_loop(feature3ds, feature2ds, move_flags, seed, seq=0, absolute_reference=1.7, window_size=5,
                 max_iterations=RANSAC_ITERS, thr=RANSAC_THR, record=False):
    """Driver gating of src/main_offline.py:57-88 (a20) around the per-frame estimator."""
    st = TemporalState(window_size)
    scales = [0]
    recs = []
    for f, mv in enumerate(move_flags):
        if not mv:
            scales.append(0)
            recs.append(None)
            continue
        f3, f2 = feature3ds[f], feature2ds[f]
        if f3.shape[0] > MIN_FEATURES:
            rec = frame_raw_scale(f3, f2, seed, f, seq, absolute_reference, max_iterations, thr)
            scales.append(st.step(rec["raw_scale"], rec["updated"]))
            recs.append(rec if record else None)
        else:
            scales.append(scales[-1])
            recs.append(None)
    return np.asarray(scales[1:], dtype=np.float64), recs
