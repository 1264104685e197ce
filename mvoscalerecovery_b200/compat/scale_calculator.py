"""API surface of the reference's older estimator (src/scale_calculator.py).  The live estimator constructs one
(rescale.py:31) but reaches it only through commented-out calls; the methods below are the host-side pieces other code
can still call.  The histogram / mode / skewness analysis and the plotting helpers (scale_calculator.py:294-364,428-600)
are not provided."""
from collections import deque

import numpy as np

from estimate_road_norm import *              # noqa: F401,F403


def bool2id(flag):
    return np.nonzero(np.asarray(flag))[0]


class ScaleEstimator:
    def __init__(self, absolute_reference, window_size=6, vanish=185, focus=718):
        self.absolute_reference = absolute_reference
        self.camera_pitch = -0.5 * np.pi / 180
        self.scale = None
        self.inliers = None
        self.scale_queue = deque()
        self.motion_queue = deque()
        self.window_size = window_size
        self.vanish = vanish
        self.focus = focus
        self.b_matrix = np.ones((3, 1), float)
        self.all_features = []
        self.correct_distance_features = []
        self.flat_features = []
        self.all_feature = []
        self.correct_distance_feature = []
        self.flat_feature = []
        self.flat_feature_2d = []
        self.img = None

    def initial_estimation(self, motion_t):
        """Pitch in degrees from the unit translation (scale_calculator.py:41-46)."""
        motion_t = np.asarray(motion_t, dtype=float)
        self.motion_queue.append(motion_t.reshape(-1))
        return np.arcsin(motion_t[1]) * 180 / np.pi

    def check_distance(self, feature3d):
        f3 = np.asarray(feature3d)
        zz = f3[:, 2] * (f3[:, 2] + 1)
        return ((zz - np.abs(self.focus * f3[:, 0])) < 0) | ((zz - np.abs(self.focus * f3[:, 1])) < 0)

    def check_triangle(self, v, d):
        a = (v[0] - v[1]) * (d[0] - d[1]) > 0
        b = (v[0] - v[2]) * (d[0] - d[2]) > 0
        c = (v[1] - v[2]) * (d[1] - d[2]) > 0
        return [bool(a or b), bool(a or b or c), bool(c)]

    def find_outliers(self, feature3d, feature2d, triangle_ids):
        f3, f2, tri = np.asarray(feature3d), np.asarray(feature2d), np.asarray(triangle_ids)
        out = np.ones(f3.shape[0])
        if tri.size:
            v, d = f2[tri, 1], f3[tri, 2]
            a = (v[:, 0] - v[:, 1]) * (d[:, 0] - d[:, 1]) > 0
            b = (v[:, 0] - v[:, 2]) * (d[:, 0] - d[:, 2]) > 0
            c = (v[:, 1] - v[:, 2]) * (d[:, 1] - d[:, 2]) > 0
            np.subtract.at(out, tri[np.stack([a | b, a | b | c, c], 1)], 1.0)
        return out

    def feature_remap(self, feature3d):
        """Rotate (y,z) by camera_pitch IN PLACE, as the reference does (scale_calculator.py:390-394)."""
        cp, sp = np.cos(self.camera_pitch), np.sin(self.camera_pitch)
        y = feature3d[:, 1] * cp - feature3d[:, 2] * sp
        z = feature3d[:, 1] * sp + feature3d[:, 2] * cp
        feature3d[:, 1] = y
        feature3d[:, 2] = z

    def scale_filtering(self, scale):
        self.scale_queue.append(scale)
        if len(self.scale_queue) > self.window_size:
            self.scale_queue.popleft()
        return np.median(self.scale_queue)

    def road_model_calculation_ransac(self, feature3d):
        """30-iteration plane RANSAC, inliers at 0.01 (scale_calculator.py:366-384): returns (height, pitch, inlier std)."""
        pts = np.asarray(feature3d, dtype=float)
        m, _ = get_pitch_ransac(pts, 30, 0.01)
        m = np.asarray(m, dtype=float)
        self.inliers = get_inliers(m, pts, 0.01)
        n, h_bar = m[:3], -m[3]
        if n[1] < 0:
            n, h_bar = -n, -h_bar
        nn = np.linalg.norm(n)
        return h_bar / nn, np.arcsin(n[1] / nn), float(np.std(pts[self.inliers, 1])) if self.inliers.any() else 0.0
