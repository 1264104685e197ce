"""Writes tests/golden/compat_api.npz: outputs of the REFERENCE's own helper functions (src/graph.py, src/rescale.py,
src/estimate_road_norm.py) on small seeded inputs, for the API-surface tests of mvoscalerecovery_b200/compat.
Run in the build container only (needs /root/reference):  python tests/golden/make_compat_golden.py"""
import contextlib
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness as H          # noqa: E402
from scipy.spatial import Delaunay           # noqa: E402


def main():
    ns = H.load_reference(seed=1)
    rng = np.random.default_rng(2026)
    n = 300
    f2 = np.stack([rng.uniform(0, 1241, n), rng.uniform(186, 376, n)], 1).astype(np.float32).astype(np.float64)
    z = 1.7 * 718.856 / (f2[:, 1] - 185.2157) * (1 + 0.02 * rng.standard_normal(n))
    z[rng.random(n) < 0.2] *= rng.uniform(0.5, 0.9)
    f3 = np.stack([(f2[:, 0] - 607.1928) * z / 718.856, (f2[:, 1] - 185.2157) * z / 718.856, z], 1).astype(np.float32).astype(np.float64)
    tri = H.canonicalise(Delaunay(f2).simplices)
    out = dict(f2=f2, f3=f3, tri=tri)
    with contextlib.redirect_stdout(io.StringIO()):
        est = ns.rescale.ScaleEstimator(1.7, 5)
        out["keep"] = ns.graph.GraphChecker([[3, 1], [2, 2], [2, 2], [0, 4]]).find_inliers(f3, f2, tri)
        out["outliers"] = est.find_outliers(f3, f2, tri)
        ids, hl = est.flat_selection(f3, tri)
        out["flat_ids"] = np.asarray(ids, np.int64); out["flat_heights"] = np.asarray(hl); out["height_level"] = np.float64(est.height_level)
        out["tp"] = ns.graph.triangle([[3, 1], [2, 2], [2, 2], [0, 4]])
        out["prob_012_211"] = np.asarray(ns.graph.get_probability([0, 1, 2], [2, 1, 1], out["tp"]))
        pts = f3[:40]
        m = ns.ern.estimate(pts[:3])
        out["plane"] = m * np.sign(m[1])
        out["inl"] = ns.ern.get_inliers(m, pts, 0.05)
        ts = rng.standard_normal((12, 3)) * 0.05 + np.array([0.0, -0.02, 1.0])
        out["ts"] = ts
        out["pitch"] = np.float64(ns.ern.get_pitch(ts)); out["pitch_svd"] = np.float64(ns.ern.get_pitch_svd(ts))
        out["norm_svd"] = np.asarray(ns.ern.get_norm_svd(ts)).reshape(-1)
        out["check_triangle"] = np.array([est.check_triangle([0., 1., 2.], [2., 1., 1.]), est.check_triangle([3., 1., 2.], [1., 2., 3.])])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "compat_api.npz"), **out)
    print("wrote compat_api.npz", {k: np.asarray(v).shape for k, v in out.items()})


if __name__ == "__main__":
    main()
