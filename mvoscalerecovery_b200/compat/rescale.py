"""Drop-in replacement of the reference's src/rescale.py: ``ScaleEstimator`` with the constructor, attributes and
method signatures of rescale.py:22-195, running the per-frame path on the GPU through libmvosr.so.

``scale_calculation(feature3d, feature2d, img=None) -> (scale, 1)`` is what src/main.py:113 and
src/main_offline.py:75 call once per frame: ROI cut, Delaunay #1, graph check, Delaunay #2, per-triangle gates,
plane RANSAC, height and raw scale run in ONE launch of the fused frame kernel (a batch of one frame); the slew limiter
and the median of the last ``window_size`` states (rescale.py:168-178) -- the estimator's temporal state -- are kept
here exactly as the reference keeps them.  There is no CPU fallback: without a CUDA device the call raises.

Differences a caller can observe, all documented in DESIGN.md:
  * RANSAC draws come from a Philox stream (``seed`` attribute, frame counter) instead of OS entropy, so runs repeat;
  * where the reference raises (QhullError on < 3 ROI features, LinAlgError, np.matrix(None)) the state is held;
  * nothing is printed.
"""
from collections import deque
import math                                   # noqa: F401  (the reference module exposes these names)

import numpy as np
from scipy.spatial import Delaunay            # noqa: F401

from estimate_road_norm import *              # noqa: F401,F403
import scale_calculator as sc
import graph
import param

import _gpu


def _engine(absolute_reference, vanish):
    return _gpu.engine(absolute_reference, vanish)


class ScaleEstimator:
    def __init__(self, absolute_reference, window_size=6):
        self.absolute_reference = absolute_reference
        self.camera_pitch = 0
        self.scale = 1
        self.inliers = None
        self.scale_queue = deque()
        self.window_size = window_size
        self.vanish = 185
        self.sc = sc.ScaleEstimator(absolute_reference, window_size)
        self.gc = graph.GraphChecker([[3, 1], [2, 2], [2, 2], [0, 4]])
        self.gs = graph.GraphGrow()
        self.img_w = param.img_w
        self.img_h = param.img_h
        self.height_level = float("nan")
        # additions: the hypothesis stream of this estimator (reference: random.seed(None) on every call)
        self.seed = 0
        self.sequence_id = 0
        self.frame_index = 0
        self.last_status = 0

    def initial_estimation(self, motion_matrix):
        return 0

    # ---- API-surface helpers (not on the mains' path) ------------------------------------------------
    def check_triangle(self, v, d):
        """[outlier flag of vertex 0, 1, 2] as the reference computes it (rescale.py:45-61)."""
        a = (v[0] - v[1]) * (d[0] - d[1]) > 0
        b = (v[0] - v[2]) * (d[0] - d[2]) > 0
        c = (v[1] - v[2]) * (d[1] - d[2]) > 0
        return [bool(a or b), bool(a or b or c), bool(c)]

    def find_outliers(self, feature3d, feature2d, triangle_ids):
        """1 - (number of triangles flagging the vertex) per feature (rescale.py:63-73)."""
        f3, f2, tri = np.asarray(feature3d), np.asarray(feature2d), np.asarray(triangle_ids)
        out = np.ones(f3.shape[0])
        if tri.size:
            v, d = f2[tri, 1], f3[tri, 2]
            a = (v[:, 0] - v[:, 1]) * (d[:, 0] - d[:, 1]) > 0
            b = (v[:, 0] - v[:, 2]) * (d[:, 0] - d[:, 2]) > 0
            c = (v[:, 1] - v[:, 2]) * (d[:, 1] - d[:, 2]) > 0
            flags = np.stack([a | b, a | b | c, c], 1)
            # the reference's fancy-indexed "-=" counts a vertex once per triangle
            np.subtract.at(out, tri[flags], 1.0)
        return out

    def flat_selection(self, feature3d, triangle_ids):
        """Per-triangle plane n = P^-1 1, pitch and height gates for caller-supplied triangles (rescale.py:75-102).
        Returns (vertex list of the valid triangles with multiplicity, heights of the loose triangles)."""
        f3, tri = np.asarray(feature3d, dtype=float), np.asarray(triangle_ids)
        p0, e1, e2 = f3[tri[:, 0]], f3[tri[:, 1]] - f3[tri[:, 0]], f3[tri[:, 2]] - f3[tri[:, 0]]
        c = np.cross(e1, e2)
        det = np.einsum("ij,ij->i", p0, c)
        if np.any(det == 0):
            raise np.linalg.LinAlgError("Singular matrix")
        clen = np.linalg.norm(c, axis=1)
        heights = np.abs(det) / clen
        pitch = np.degrees(np.arcsin(-np.sign(det) * c[:, 1] / clen))
        loose, tight = pitch < -80, pitch < -85
        self.height_level = 0.9 * np.median(heights[loose]) if loose.any() else float("nan")
        valid = tight & (heights > self.height_level)
        return list(tri[valid].reshape(-1)), heights[loose]

    # ---- the GPU path ----------------------------------------------------------------------------
    def _launch(self, feature3d, feature2d, debug):
        import torch
        eng = _engine(self.absolute_reference, self.vanish)
        f3 = np.ascontiguousarray(feature3d, dtype=np.float64).reshape(-1, 3)
        f2 = np.ascontiguousarray(feature2d, dtype=np.float64).reshape(-1, 2)
        n = f3.shape[0]
        dev = eng.device
        off = torch.tensor([0, n], dtype=torch.int32, device=dev)
        out = eng.scale_frames_f64(off, torch.from_numpy(f3).to(dev), torch.from_numpy(f2).to(dev), max(n, 1),      # device copies: the
                                   frame_index0=self.frame_index, seq_id=self.sequence_id, seed=self.seed,            # caller's arrays are
                                   stats=True, debug=debug)                                                           # never modified
        torch.cuda.synchronize(dev)
        return eng, out

    def feature_selection(self, feature3d, feature2d):
        """(point_selected (N_sel,3), heights of the loose triangles) as rescale.py:113-148 returns them."""
        from mvoscalerecovery_b200.batch import stats_to_numpy
        _, out = self._launch(feature3d, feature2d, debug=True)
        st = stats_to_numpy(out["stats"])[0]
        dbg = out["debug"]
        f3 = np.asarray(feature3d, dtype=np.float64).reshape(-1, 3)
        roi = f3[np.asarray(feature2d)[:, 1] > self.vanish]
        keep = dbg["keep"].cpu().numpy()[: roi.shape[0]].astype(bool)
        kept = roi[keep] if int(st["n_kept"]) > 10 else roi
        data_id = dbg["data_id"].cpu().numpy()[: 3 * int(st["n_valid"])]
        flags = dbg["tri_flags"].cpu().numpy()[: int(st["n_tri"])]
        heights = dbg["tri_height"].cpu().numpy()[: int(st["n_tri"])]
        self.height_level = float(st["height_level"])
        return kept[data_id], heights[(flags & 1) != 0]

    def _apply_state(self, raw_scale, updated):
        if updated:
            if raw_scale - self.scale > 0.3:
                self.scale += 0.3
            elif raw_scale - self.scale < -0.3:
                self.scale -= 0.3
            else:
                self.scale = raw_scale
        self.scale_queue.append(self.scale)
        if len(self.scale_queue) > self.window_size:
            self.scale_queue.popleft()
        return np.median(self.scale_queue), 1

    def scale_calculation_ransac(self, point_selected):
        """Plane RANSAC over an explicit vertex list + the temporal state (rescale.py:151-178).  Host-side helper for
        callers that bring their own selection; scale_calculation does the same inside the frame kernel."""
        pts = np.asarray(point_selected, dtype=float)
        raw, updated = float("nan"), False
        if pts.shape[0] >= 12:
            m, _ = get_pitch_ransac(pts, 100, 0.005)
            m = np.asarray(m, dtype=float)
            n, h_bar = m[:3], -m[3]
            if n[1] < 0:
                n, h_bar = -n, -h_bar
            raw = self.absolute_reference / (h_bar / np.linalg.norm(n))
            updated = True
        return self._apply_state(raw, updated)

    def scale_calculation_static_tri(self, heights):
        if len(heights) > 12:
            scale_norm, _, _ = self.sc.road_model_calculation_static_tri(heights)
            self.scale = scale_norm * self.absolute_reference
            return self.scale, 0
        return self.scale, 0

    def scale_calculation(self, feature3d, feature2d, img=None):
        """One C-ABI call per frame (mvosr_scale_frame_host_f64): the caller's float64 arrays are copied as they are -- the ROI cut,
        the votes, the gates and the RANSAC evaluate the float64 values, as the reference does --, one launch, one copy back."""
        from mvoscalerecovery_b200 import _native as N
        eng = _engine(self.absolute_reference, self.vanish)
        f3 = np.ascontiguousarray(feature3d, dtype=np.float64).reshape(-1, 3)        # no copy when the caller passes float64 C arrays
        f2 = np.ascontiguousarray(feature2d, dtype=np.float64).reshape(-1, 2)
        raw, status, _, st = eng.scale_frame_host_f64(f3, f2, frame_index=self.frame_index, seq_id=self.sequence_id, seed=self.seed)
        self.height_level = float(st.height_level)
        self.last_status = status
        self.frame_index += 1
        return self._apply_state(raw, bool(status & N.ST_UPDATED))
