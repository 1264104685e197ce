"""Writes tests/golden/bucket.npz: outputs of the reference's own bucket() (src/detector.py:65-95, imported unmodified from
/root/reference/src with stdout silenced) on seeded KITTI-shaped feature sets, numpy's global RNG seeded.  Which members of a cell
survive depends on that RNG and is not reproducible by design; the golden pins the cells visited, their order and the number of
survivors per cell.  Run in the build container:  python tests/golden/make_bucket_golden.py"""
import contextlib
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, "/root/reference/src")
import detector          # noqa: E402  (the reference, unmodified)


def main():
    rng = np.random.default_rng(42)
    out = {}
    for k, (n, bs, dens) in enumerate(((1500, 30, 2), (4000, 30, 2), (300, 50, 1), (900, 20, 3), (40, 30, 2))):
        f = np.stack([rng.uniform(0, 1241, n), rng.uniform(0, 376, n) ** (1.0 if k % 2 else 0.8)], 1).astype(np.float32)
        f[: n // 10] = np.floor(f[: n // 10])                      # integer pixels: cell borders
        np.random.seed(7 + k)
        with contextlib.redirect_stdout(io.StringIO()):
            kept = detector.bucket(f, bs, dens)
        out["f%d" % k] = f; out["kept%d" % k] = kept; out["par%d" % k] = np.array([bs, dens])
    out["n_sets"] = 5
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "bucket.npz"), **out)
    print({k: v.shape for k, v in out.items() if hasattr(v, "shape")})


if __name__ == "__main__":
    main()
