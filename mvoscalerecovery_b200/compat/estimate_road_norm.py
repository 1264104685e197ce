"""API surface of the reference's src/estimate_road_norm.py (plane / line models, pitch helpers).  Small host-side
numpy helpers with the reference's signatures and conventions.  The plane RANSAC runs on the GPU: per frame it is stage 4
of the CUDA frame kernel (rescale.ScaleEstimator.scale_calculation), and ``get_pitch_ransac`` on an explicit point list is
mvosr_ransac_planes (no CPU fallback).  Its draws come from the Philox position stream: key
``ransac_seed``, one "frame" counter per call, so a process repeats its results (the reference re-seeds from OS entropy
on every call, src/thirdparty/Ransac/ransac.py:6)."""
import math
import sys

import numpy as np
from scipy.spatial import Delaunay          # re-exported: the reference's modules rely on this module's namespace

try:                                         # the reference re-exports cv2 through `import *` (src/estimate_road_norm.py:3)
    import cv2
except Exception:                            # pragma: no cover
    cv2 = None
from thirdparty.Ransac.ransac import *      # noqa: F401,F403  (run_ransac, random)
from thirdparty.Ransac.ransac import run_ransac

camera_focus = 718.856
camera_cx = 607.1928
camera_cy = 185.2157


def augment(xyzs):
    """[p 1] rows (estimate_road_norm.py:8-11)."""
    p = np.asarray(xyzs, dtype=float)
    return np.hstack([p, np.ones((p.shape[0], 1))])


def estimate(xyzs):
    """Plane through the first three points: unit-4-norm null vector of the 3x4 matrix [p 1] (:13-15)."""
    return np.linalg.svd(augment(xyzs[:3]))[-1][-1, :]


def is_inlier(coeffs, xyz, threshold):
    """|a x + b y + c z + d| < threshold, algebraic residual (:17-18)."""
    return np.abs(np.asarray(coeffs).dot(augment([xyz]).T)) < threshold


def _down_normal(camera_motion_ts):
    u = np.linalg.svd(np.asarray(camera_motion_ts, dtype=float).T, full_matrices=True)[0][:, 2]
    return -u if u[1] < 0 else u


def get_norm_svd(camera_motion_ts):
    """Least-variance direction of the stacked translations, y >= 0, as a 1x3 matrix (:20-26)."""
    return np.matrix(_down_normal(camera_motion_ts))


def get_pitch_svd(camera_motion_ts):
    """asin(n_y / |n|^2) of that direction (:28-37; the reference divides by the squared norm)."""
    n = _down_normal(camera_motion_ts)
    return math.asin(n[1] / float(n.dot(n)))


def augment_line(xys):
    p = np.asarray(xys, dtype=float)
    return np.hstack([p, np.ones((p.shape[0], 1))])


def estimate_line(xys):
    return np.linalg.svd(augment_line(xys[:2]))[-1][-1, :]


def is_inlier_line(coeffs, xy, threshold):
    return np.abs(np.asarray(coeffs).dot(augment_line([xy]).T)) < threshold


def get_pitch(camera_motion_ts):
    """asin(-sum(t)_y / |sum(t)|^2) (:52-58)."""
    s = np.sum(np.asarray(camera_motion_ts, dtype=float), 0)
    return math.asin(-s[1] / float(s.dot(s)))


def get_pitch_line_ransac(road_points, max_iterations, threshold):
    goal = road_points.shape[0] * 0.8
    return run_ransac(road_points, estimate_line, lambda m, p: is_inlier_line(m, p, threshold), 2, goal, max_iterations)


ransac_seed = 0            # Philox key of get_pitch_ransac
ransac_calls = 0           # number of get_pitch_ransac calls so far = the stream's "frame" counter of the next call


def get_pitch_ransac(road_points, max_iterations, threshold):
    """3-point plane RANSAC, goal 0.8 N, stop at the first count above the goal (:66-70).  Returns (model (4,), inlier
    count); the model is None when no hypothesis had an inlier, as run_ransac returns it."""
    global ransac_calls
    import _gpu
    frame = ransac_calls
    ransac_calls += 1
    m, ic, best, _ = _gpu.ransac_plane(road_points, max_iterations, threshold, seed=ransac_seed, frame=frame)
    return (m if best >= 0 else None), ic


def get_inliers(parameter, data, threshold):
    """|n . x + d| < threshold for every row of data (:71-78)."""
    m = np.asarray(parameter, dtype=float).reshape(-1)
    d = np.asarray(data, dtype=float)
    return np.abs(d.dot(m[:-1]) + m[-1]) < threshold
