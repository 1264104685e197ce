"""Drop-in for the reference's older estimator, src/scale_calculator.py: same class, constructor, attributes, method
names, argument meaning and return values (file:line citations below are into the reference tree).

The live estimator (rescale.py) constructs one of these (rescale.py:31) and reaches it through
``scale_calculation_static_tri``; ``scale_calculation`` is the older complete pipeline (ROI -> Delaunay -> vote-based
outlier rejection -> Delaunay -> per-triangle pitch/height gates -> histogram-mode road height -> median filter).
Division of labour here:
  * both Delaunay triangulations, the per-triangle planes and the depth-order votes run on the GPU through libmvosr.so
    (``_gpu.py``: mvosr_delaunay_frames / mvosr_triangle_planes / mvosr_triangle_votes) -- no scipy/Qhull, no CPU fallback;
  * the histogram / mode / skewness analysis (scale_calculator.py:294-364,428-497) is a few dozen scalar operations per
    frame and stays host code;
  * the two belief-propagation variants (find_reliability_by_graph, feature_selection_by_tri_graph) are strictly
    sequential in-place recurrences over the mesh and stay host loops;
  * plotting (distribution, plot_distribution) needs matplotlib and is GUI code: available only if matplotlib is installed.
Nothing is printed.
"""
from collections import deque

import numpy as np

from estimate_road_norm import *              # noqa: F401,F403  (the reference re-exports np, cv2, Delaunay, math, run_ransac ...)
from estimate_road_norm import get_pitch_ransac, get_inliers
import _gpu

_OBSERVATION = np.array([[0.33, 0.33, 0.33], [0.03, 0.07, 0.90], [0.90, 0.07, 0.03], [0.05, 0.9, 0.05]])


def bool2id(flag):
    """Indices of the set entries (scale_calculator.py:604-607)."""
    return np.nonzero(np.asarray(flag))[0]


def draw_feature(img, feature, color=(255, 255, 0)):
    """Filled circles at the feature pixels (scale_calculator.py:609-611)."""
    for u, v in np.asarray(feature)[:, :2]:
        cv2.circle(img, (int(u), int(v)), 3, color, -1)     # noqa: F405


def _edge_neighbours(triangle_ids):
    """Triangle adjacency in the reference's insertion order (triangle2region_graph, :56-82): triangle i is linked to the
    earlier triangle that shares edge (a,b), then (a,c), then (b,c)."""
    tri = np.asarray(triangle_ids)
    graph = [[] for _ in range(tri.shape[0])]
    first = {}
    for i, (a, b, c) in enumerate(tri.tolist()):
        for e in ((a, b), (a, c), (b, c)):
            key = e if e[0] < e[1] else (e[1], e[0])
            j = first.get(key)
            if j is None:
                first[key] = i
            else:
                graph[i].append(j)
                graph[j].append(i)
    return graph


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

    # ------------------------------------------------------------------ small helpers
    def initial_estimation(self, motion_t):
        """Pitch in degrees from the unit translation (:41-46)."""
        motion_t = np.asarray(motion_t, dtype=float)
        self.motion_queue.append(motion_t.reshape(-1))
        return np.arcsin(motion_t[1]) * 180 / np.pi

    def check_distance(self, feature3d):
        """z (z + 1) < focus |x| or focus |y| (:49-54)."""
        f3 = np.asarray(feature3d)
        zz = f3[:, 2] * (f3[:, 2] + 1)
        return ((zz - np.abs(self.focus * f3[:, 0])) < 0) | ((zz - np.abs(self.focus * f3[:, 1])) < 0)

    def triangle2region_graph(self, triangle_ids):
        return _edge_neighbours(triangle_ids)

    def triangle2graph(self, triangle_ids):
        """Vertex adjacency towards the larger index, in order of first appearance (:86-99)."""
        tri = np.sort(np.asarray(triangle_ids), axis=1)
        graph = [[] for _ in range(int(tri.max()) + 1)]
        for a, b, c in tri.tolist():
            for p, q in ((a, b), (a, c), (b, c)):
                if q not in graph[p]:
                    graph[p].append(q)
        return graph

    def check_triangle(self, v, d):
        """Outlier flags of the three vertices (:105-119)."""
        a = (v[0] - v[1]) * (d[0] - d[1]) > 0
        b = (v[0] - v[2]) * (d[0] - d[2]) > 0
        c = (v[1] - v[2]) * (d[1] - d[2]) > 0
        return np.array([a or b, a or b or c, bool(c)])

    def check_depth(self, v, d):
        return bool((v[0] - v[1]) * (d[0] - d[1]) > 0)

    def compare(self, a, b, threshold=0.1):
        return -1 if a - b < -threshold else (1 if a - b > threshold else 0)

    # ------------------------------------------------------------------ outlier rejection
    def find_reliability_by_graph(self, feature3d, feature2d, triangle_ids):
        """Pairwise belief update over the mesh edges, in place and in edge order (:127-149): reliability > 0.8 survives."""
        f3, f2 = np.asarray(feature3d), np.asarray(feature2d)
        rel = np.full(f3.shape[0], 0.8)
        v, d = f2[:, 1], f3[:, 2]
        for i, nbrs in enumerate(self.triangle2graph(triangle_ids)):
            for j in nbrs:
                ri, rj = rel[i], rel[j]
                both, only_j, only_i, none = ri * rj, (1 - ri) * rj, (1 - rj) * ri, (1 - ri) * (1 - rj)
                wrong_order = (v[i] - v[j]) * (d[i] - d[j]) > 0
                num_both = 0.0 if wrong_order else both
                z = num_both + 0.25 * (only_j + only_i) + 0.5 * none
                rel[i] = (num_both + 0.25 * only_i) / z
                rel[j] = (num_both + 0.25 * only_j) / z
        return rel > 0.8

    def find_outliers(self, feature3d, feature2d, triangle_ids):
        """Vote count 1 + (#triangles not flagging) - (#triangles flagging) >= 0 (:151-167).  Votes on the GPU."""
        flagged, incident = _gpu.triangle_votes(triangle_ids, np.asarray(feature2d)[:, 1], np.asarray(feature3d)[:, 2])
        return (1.0 + incident - 2.0 * flagged) >= 0

    # ------------------------------------------------------------------ per-triangle selection
    def _triangle_geometry(self, feature3d, triangle_ids):
        """pitch (deg) of n = P^-1 1 and the mean Y of every triangle (:228-237); planes on the GPU."""
        normal, _, mean_y = _gpu.triangle_planes(triangle_ids, feature3d)
        pitch_deg = np.arcsin(-normal[:, 1] / np.sqrt(np.sum(normal * normal, 1))) * 180 / np.pi
        return pitch_deg, mean_y

    def feature_selection_by_tri_graph(self, feature3d, triangle_ids):
        """Road probability per triangle refined over its neighbours' height ordering (:177-223)."""
        tri = np.asarray(triangle_ids)
        graph = _edge_neighbours(tri)
        pitch_deg, heights = self._triangle_geometry(feature3d, tri)
        p_road = np.maximum((-70 - pitch_deg) / 20 - 0.2, 0)
        flat = pitch_deg < -80
        for t in bool2id(flat):
            pa, ha = p_road[t], heights[t]
            for nb in graph[t]:
                pc = p_road[nb]
                col = self.compare(heights[nb], ha) + 1
                joint = np.array([(1 - pa) * (1 - pc), (1 - pa) * pc, pa * (1 - pc), pa * pc])
                pa = _OBSERVATION[2:4, col] @ joint[2:4] / (_OBSERVATION[:, col] @ joint)
            p_road[t] = pa
        self.height_level = np.mean(heights[~flat])
        return np.unique(tri[p_road > 0.5].reshape(-1))

    def feature_selection_by_tri(self, feature3d, triangle_ids):
        """Vertices of the triangles with pitch < -80 deg lying below the mean height of the others (:225-248)."""
        tri = np.asarray(triangle_ids)
        pitch_deg, heights = self._triangle_geometry(feature3d, tri)
        flat = pitch_deg < -80
        self.height_level = np.mean(heights[~flat])
        return np.unique(tri[flat & (heights > self.height_level)].reshape(-1))

    def feature_selection(self, feature3d, feature2d):
        """ROI cut, Delaunay, vote rejection, Delaunay again, flat-triangle selection (:250-279)."""
        f3, f2 = np.asarray(feature3d), np.asarray(feature2d)
        low = f2[:, 1] > self.vanish
        f2, f3 = f2[low], f3[low]
        valid = self.find_outliers(f3, f2, _gpu.delaunay(f2))
        if valid.shape[0] <= 3:                    # (the reference tests the length of the mask, not its sum)
            return None
        f2, f3 = f2[valid], f3[valid]
        selected = self.feature_selection_by_tri(f3, _gpu.delaunay(f2))
        if len(selected) == 0:
            return None
        self.flat_feature_2d = f2[selected]
        return f3[selected]

    # ------------------------------------------------------------------ road model
    def road_model_calculation(self, feature3d):
        return self.road_model_calculation_static(feature3d)

    def remove_single(self, feature3d, dis, bins):
        """Drop the features that sit alone in a 0.1-wide height bin (:284-293)."""
        f3 = np.asarray(feature3d)
        edges = np.asarray(bins)[1:][np.asarray(dis) == 1]
        if np.sum(edges) > 0:
            y = f3[:, 1]
            alone = (y >= edges[0] - 0.1) & (y <= edges[0])
            for e in edges[1:]:
                alone |= (y > e - 0.1) & (y <= e)
            f3 = f3[~alone, :]
        return f3

    @staticmethod
    def _mode_span(group):
        """int(10 x) of the first and last bin edge of a mode group (:306-307,340-341)."""
        g = np.asarray(group, dtype=float).reshape(-1)
        return int(g[0] * 10), int(g[-1] * 10)

    def road_model_calculation_static_tri(self, heights):
        """Mode of the histogram of 1/height over 0.1-wide bins, else the median (:294-322).  Returns (value, 0, 1)."""
        inv = 1 / np.asarray(heights, dtype=float)
        dis, bins = np.histogram(inv, bins=np.arange(20) * 0.1)
        dis[dis == 1] = 0
        modes = self.check_mode(dis, bins)
        if len(modes) == 0:
            return np.median(inv), 0, 1
        left, right = self._mode_span(modes[0])
        return ((left + right) / 2) / 10, 0, 1

    def road_model_calculation_static(self, feature3d):
        """Histogram-mode road height from the Y of the selected features (:324-364).  Returns (height, 0, 1)."""
        f3 = np.asarray(feature3d)
        dis, bins = np.histogram(f3[:, 1], bins=np.arange(170) * 0.1)
        f3 = self.remove_single(f3, dis, bins)
        dis[dis == 1] = 0
        modes = self.check_mode(dis, bins)
        if len(modes) == 0:
            return (np.median(f3[:, 1]) if f3.shape[0] > 0 else self.height_level), 0, 1
        valleys = self.check_reverse_mode(dis)
        mode_left, mode_right = self._mode_span(modes[-1])
        mode = (mode_left + mode_right) / 2
        left = bins[1:mode_left + 1][valleys[:mode_left]][-1]           # noqa: F841  (evaluated for its IndexError, as the reference does)
        right = bins[mode_right + 1:][valleys[mode_right:]][0]
        if self.check_skewness(f3[:, 1], mode=mode / 10) > 0.3:
            return right, 0, 1
        return mode / 10, 0, 1

    def road_model_calculation_ransac(self, feature3d):
        """30-iteration plane RANSAC at 0.005, inliers at 0.01 (:366-384).  Returns (height, pitch, inlier points)."""
        pts = np.asarray(feature3d)
        m, _ = get_pitch_ransac(np.array(pts), 30, 0.005)
        inliers = pts[get_inliers(m, pts, 0.01), :]
        m = np.array(m)
        normal, h_bar = m[:-1], -m[-1]
        if normal[1] < 0:
            normal, h_bar = -normal, -h_bar
        length = np.sqrt(np.sum(normal * normal))
        return h_bar / length, np.arcsin(-normal[1] / length), inliers

    def distribution(self, feature3d):
        import matplotlib.pyplot as plt
        dis, _ = np.histogram(np.asarray(feature3d)[:, 1], bins=np.arange(100) * 0.1)
        plt.plot(dis)

    def feature_remap(self, feature3d):
        """Rotate (y,z) by camera_pitch IN PLACE, as the reference does (:390-394)."""
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

    def _scale_from_selection(self, point_selected):
        std = 100
        if point_selected is not None:
            height, _, std = self.road_model_calculation(point_selected)
            scale = self.absolute_reference / height
        else:
            scale = self.absolute_reference / self.height_level
        return self.scale_filtering(scale), std

    def scale_calculation_static(self, point_selected):
        self.feature_remap(point_selected)
        return self._scale_from_selection(point_selected)

    def scale_calculation(self, feature3d, feature2d, img=None):
        """The older per-frame entry (:411-423): remaps feature3d IN PLACE, selects, histogram-mode height, median filter."""
        self.feature_remap(feature3d)
        self.flat_feature = self.feature_selection(feature3d, feature2d)
        return self._scale_from_selection(self.flat_feature)

    # ------------------------------------------------------------------ histogram analysis
    def check_reverse_mode(self, dis_data):
        """Local minima of the histogram (:428-443)."""
        d = np.asarray(dis_data)
        flag = np.zeros(d.shape[0], dtype=bool)
        flag[0] = d[0] == d.min()
        flag[-1] = d[-1] == d.min()
        mid, lft, rgt = d[1:-1], d[:-2], d[2:]
        flag[1:-1] = (mid <= lft) & (mid <= rgt) & ~((mid == rgt) & (mid == lft))
        return flag

    def check_mode(self, dis_data, bins):
        """Groups of adjacent local maxima (>= 2 counts, >= a third of the peak), as lists of upper bin edges (:446-483)."""
        d = np.asarray(dis_data)
        peak = d.max()
        if peak <= 2:
            return []
        flag = np.zeros(d.shape[0], dtype=bool)
        flag[0] = d[0] == peak
        flag[-1] = d[-1] == peak
        mid = d[1:-1]
        flag[1:-1] = (mid >= d[:-2]) & (mid >= d[2:]) & (mid >= 0.33 * peak) & (mid >= 2)
        edges = np.asarray(bins)[1:][flag]
        if edges.size == 1:
            return [[edges]]
        groups, cur = [], []
        for e in edges:
            if cur and not (e - cur[-1] < 0.11):
                groups.append(cur)
                cur = []
            cur.append(e)
        if cur:
            groups.append(cur)
        return groups

    def check_skewness(self, data, mode=None, method='p1'):
        """Pearson skewness: (mean - mode)/std ('p1') or 3 (mean - median)/std ('p2') (:486-497)."""
        data = np.asarray(data)
        if method == 'p2':
            return 3 * (np.mean(data) - np.median(data)) / np.std(data)
        if mode is None:
            dis, bins = np.histogram(data, bins=np.arange(170) * 0.1)
            dis[dis == 1] = 0
            mode = np.mean(bins[1:][dis == np.max(dis)])
        return (np.mean(data) - mode) / np.std(data)

    def skewness_analysis(self):
        return self.check_skewness(self.flat_feature[:, 1])

    def mode_analysis(self):
        dis, bins = np.histogram(self.flat_feature[:, 1], bins=np.arange(100) * 0.1, density=True)
        return self.check_mode(dis, bins)

    # ------------------------------------------------------------------ diagnostics (GUI)
    def plot_distribution(self, label, img, scale=1):
        """Three-panel diagnostic figure (:509-560); needs matplotlib."""
        import matplotlib.pyplot as plt
        ax = plt.subplot(221)
        if self.flat_feature is not None and len(self.flat_feature) > 0:
            dis, bins = np.histogram(np.asarray(self.flat_feature)[:, 1], bins=np.arange(50) * 0.1, density=True)
            ax.plot(bins[:-1], 0.1 * dis, 'y-*', label='Selected Features')
        ax.set_title('vertical distribution')
        ax.set_xlabel('y')
        ax = plt.subplot(222)
        for pts, style in ((self.all_feature, '.r'), (self.correct_distance_feature, '.g'), (self.flat_feature, '.y')):
            if pts is not None and len(pts) > 0:
                ax.plot(np.asarray(pts)[:, 2], -np.asarray(pts)[:, 1], style)
        ax.set_xlabel('z')
        ax.set_ylabel('-y')
        ax.set_title('feature projection')
        plt.subplot(212).imshow(self.img)
        plt.show()

    def check_full_distribution(self, feature3d, feature2d, scale, img):
        """Records the three feature populations (all / depth-consistent / flat) of one frame for plotting (:562-601)."""
        f3, f2 = np.asarray(feature3d), np.asarray(feature2d)
        low = f2[:, 1] > self.vanish
        f3, f2 = f3[low], f2[low]
        self.all_feature = f3.copy()
        self.all_features.extend(f3 * scale)
        draw_feature(img, f2, (255, 0, 0))
        valid = self.find_reliability_by_graph(f3, f2, _gpu.delaunay(f2))
        if valid.shape[0] <= 3:
            return
        f2, f3 = f2[valid], f3[valid]
        draw_feature(img, f2, (0, 255, 0))
        self.correct_distance_feature = f3.copy()
        self.correct_distance_features.extend(f3 * scale)
        selected = self.feature_selection_by_tri(f3, _gpu.delaunay(f2))
        if len(selected) > 0:
            self.flat_feature = f3[selected].copy()
            self.flat_features.extend(f3[selected] * scale)
            draw_feature(img, f2[selected], (255, 255, 0))
        self.img = img
