"""Deterministic harness around the UNMODIFIED reference -- TEST INFRASTRUCTURE.

It imports the reference's own modules from ``/root/reference/src`` (build container) or
from the staged copy ``oracle/_ref/src`` that ``make -C oracle ref`` makes of the same files
(git-ignored; it travels to the GPU box, where bench.py's CPU legs time it) under
the external shims of SURVEY.md section 8(c) and records, per frame, every
intermediate the parity tests compare.  Nothing here restates the algorithm:
values are read out of the reference's own stack frames with ``sys.setprofile``.

Shims (reference files untouched):
  * ``matplotlib`` / ``matplotlib.pyplot`` -> empty stub modules
    (imported at src/rescale.py:17, src/scale_calculator.py:17; not installed here)
  * ``np.float = float`` (src/rescale.py:76, src/scale_calculator.py:32; removed in numpy>=1.24)
  * ``thirdparty.Ransac.ransac.random`` -> ``oracle.philox.PhiloxShim``
    (the reference re-seeds from OS entropy per call, ransac.py:6 -> not reproducible)
  * ``rescale.Delaunay`` -> wrapper around the same ``scipy.spatial.Delaunay`` (Qhull)
    whose ``.simplices`` are canonicalised (rows sorted ascending, then lexsorted), because
    Qhull's simplex order is an implementation detail and the vertex list handed to
    RANSAC is "in triangle order" (src/rescale.py:101).
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

import numpy as np

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "src")     # `make -C oracle ref`: the same files, staged
REF_SRC = os.environ.get("MVOSR_REFERENCE_SRC") or ("/root/reference/src" if os.path.isfile("/root/reference/src/rescale.py") else _STAGED)

_loaded = None


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_SRC, "rescale.py"))


def canonicalise(simplices: np.ndarray) -> np.ndarray:
    s = np.sort(np.asarray(simplices, dtype=np.int64), axis=1)
    order = np.lexsort((s[:, 2], s[:, 1], s[:, 0]))
    return s[order].astype(np.int32)


def load_reference(seed: int = 0):
    """Import the reference modules once, under the shims. Returns a namespace."""
    global _loaded
    if _loaded is not None:
        _loaded.shim.seed_value = int(seed)
        return _loaded
    if not reference_available():
        raise RuntimeError("reference sources not found at %s" % REF_SRC)
    from oracle.philox import PhiloxShim
    for m in ("matplotlib", "matplotlib.pyplot"):
        if m not in sys.modules:
            sys.modules[m] = types.ModuleType(m)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if not hasattr(np, "float"):
        np.float = float            # noqa: removed alias the reference still uses
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    with contextlib.redirect_stdout(io.StringIO()):
        import rescale                                   # noqa: E402  (reference module)
        import graph                                     # noqa: E402
        import estimate_road_norm                        # noqa: E402
        import thirdparty.Ransac.ransac as ransac        # noqa: E402
        import param                                     # noqa: E402
    from scipy.spatial import Delaunay as _QhullDelaunay

    ns = types.SimpleNamespace()
    ns.rescale, ns.graph, ns.ern, ns.ransac, ns.param = rescale, graph, estimate_road_norm, ransac, param
    ns.shim = PhiloxShim(seed)
    ransac.random = ns.shim
    ns.dt_log = []

    class CanonDelaunay:
        def __init__(self, points):
            self._tri = _QhullDelaunay(points)
            self.simplices = canonicalise(self._tri.simplices)
            self.raw_simplices = self._tri.simplices
            self.coplanar = self._tri.coplanar
            ns.dt_log.append(self)

    rescale.Delaunay = CanonDelaunay
    ns.CanonDelaunay = CanonDelaunay
    _loaded = ns
    return ns


class _LocalsProbe:
    """Capture selected locals of named reference functions at their return."""

    def __init__(self, wanted):
        self.wanted = wanted          # {code name: (filename suffix, [local names])}
        self.out = {}

    def __call__(self, frame, event, arg):
        if event != "return":
            return
        code = frame.f_code
        spec = self.wanted.get(code.co_name)
        if spec is None or not code.co_filename.endswith(spec[0]):
            return
        loc = frame.f_locals
        self.out[code.co_name] = {k: loc[k] for k in spec[1] if k in loc}


_PROBES = {
    "flat_selection": ("rescale.py", ["pitch_deg", "heights", "valid_pitch_id", "valid_pitch_id_tight",
                                      "height_level", "valid_id", "valid_points_id", "normals"]),
    "feature_selection": ("rescale.py", ["valid_id", "lower_feature_ids", "data_id", "point_selected"]),
    "scale_calculation_ransac": ("rescale.py", ["m", "b", "scale", "ransac_camera_height", "pitch"]),
}


def run_frame(ns, estimator, feature3d, feature2d, frame, seq=0):
    """One ``scale_calculation`` call of the reference with every intermediate recorded."""
    ns.shim.begin_frame(frame, seq)
    ns.dt_log.clear()
    probe = _LocalsProbe(_PROBES)
    state_before = float(estimator.scale)
    sys.setprofile(probe)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            scale_out, std = estimator.scale_calculation(feature3d, feature2d)
    finally:
        sys.setprofile(None)
    fs = probe.out.get("flat_selection", {})
    fe = probe.out.get("feature_selection", {})
    sr = probe.out.get("scale_calculation_ransac", {})
    # the returned model's inlier set over the vertex list, by the reference's own is_inlier (estimate_road_norm.py:17-18;
    # threshold 0.005 as get_pitch_ransac is called, rescale.py:155) -- "RANSAC inlier index sets" of the north star
    inlier = np.zeros(0, dtype=bool)
    if "m" in sr and sr["m"] is not None and "point_selected" in fe:
        inlier = np.array([bool(ns.ern.is_inlier(sr["m"], p, 0.005)) for p in np.asarray(fe["point_selected"])], dtype=bool)
    rec = {
        "inlier": inlier,
        "roi": np.asarray(fe.get("lower_feature_ids")),
        "tri1": ns.dt_log[0].simplices,
        "coplanar1": np.asarray(ns.dt_log[0].coplanar),
        "keep": np.asarray(fe.get("valid_id")),
        "tri2": ns.dt_log[-1].simplices,
        "second_dt": len(ns.dt_log) > 1,
        "pitch_deg": np.asarray(fs.get("pitch_deg")),
        "heights": np.asarray(fs.get("heights")),
        "loose": np.asarray(fs.get("valid_pitch_id")),
        "tight": np.asarray(fs.get("valid_pitch_id_tight")),
        "valid": np.asarray(fs.get("valid_id")),
        "height_level": float(fs.get("height_level")),
        "data_id": np.asarray(fe.get("data_id"), dtype=np.int32),
        "hyp_log": np.asarray(ns.shim.log, dtype=np.int64).reshape(-1, 4),
        "model": np.asarray(sr["m"], dtype=np.float64) if "m" in sr and sr["m"] is not None else np.full(4, np.nan),
        "best_ic": int(sr["b"]) if "b" in sr else -1,
        "raw_scale": float(sr["scale"]) if "scale" in sr else np.nan,
        "height": float(sr["ransac_camera_height"]) if "ransac_camera_height" in sr else np.nan,
        "updated": "scale" in sr,
        "state_before": state_before,
        "state_after": float(estimator.scale),
        "scale_out": float(scale_out),
        "std": std,
    }
    return rec


def run_offline_loop(ns, feature3ds, feature2ds, move_flags, absolute_reference=1.7, window_size=5,
                     seq=0, record=True):
    """The loop of src/main_offline.py:57-88 (the only thing restated here is that driver
    loop, because it lives inside ``main()`` and reads ``sys.argv``); every estimator call
    goes to the unmodified reference.  Returns (scales[1:], per-frame records or None)."""
    est = ns.rescale.ScaleEstimator(absolute_reference=absolute_reference, window_size=window_size)
    scales = [0]
    recs = []
    image_id = 0
    for move_flag in move_flags:
        if not move_flag:
            scales.append(0)
            recs.append(None)
            image_id += 1
            continue
        f3 = feature3ds[image_id]
        f2 = feature2ds[image_id]
        if f3.shape[0] > ns.param.minimum_feature_for_scale:
            est.initial_estimation(np.zeros(3))
            rec = run_frame(ns, est, f3, f2, image_id, seq)
            scales.append(rec["scale_out"])
            recs.append(rec if record else None)
        else:
            scales.append(scales[-1])
            recs.append(None)
        image_id += 1
    return np.asarray(scales[1:], dtype=np.float64), recs


def reference_filter10(data):
    """``filter(data, window=10)`` of script/evaluate_scale.py:25-29, executed from the
    reference file itself (the module imports matplotlib at top level -> stubbed)."""
    load_reference()
    import importlib.util
    path = os.path.join(os.path.dirname(REF_SRC), "script", "evaluate_scale.py")
    spec = importlib.util.spec_from_file_location("_ref_evaluate_scale", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.filter(np.asarray(data, dtype=np.float64), 10)
