"""Writes tests/golden/essential.npz: outputs of oracle/five_point_plan.find_essential_philox (the CPU restatement of the
five-point RANSAC on the shared Philox stream) on the frames of pose.npz with 20 % of the correspondences replaced by seeded
mismatches, plus three edge frames (four correspondences, exactly five, six copies of one correspondence).  The golden pins
what mvosr_find_essential_frames is compared with on the GPU box, where the Python oracle would be too slow to run per test.
Run in the build container:  python tests/golden/make_essential_golden.py   (about half a minute)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import five_point_plan as PL          # noqa: E402

K = (718.856, 718.856, 607.1928, 185.2157)
HYPOTHESES, THRESHOLD, SEED, SEQ = 64, 0.5, 2024, 3


def build_inputs():
    z = np.load(os.path.join(ROOT, "tests", "golden", "pose.npz"))
    off = z["offsets"]
    rng = np.random.default_rng(5)
    cu, cv, ru, rv, truth, lens = [], [], [], [], [], []
    for f in range(len(off) - 1):
        a, e = off[f], off[f + 1]
        n = e - a
        r_u, r_v = z["ref_u"][a:e].copy(), z["ref_v"][a:e].copy()
        bad = rng.permutation(n)[: n // 5]
        r_u[bad] = rng.uniform(0, 1241, bad.size).astype(np.float32)
        r_v[bad] = rng.uniform(0, 376, bad.size).astype(np.float32)
        ok = np.ones(n, bool); ok[bad] = False
        cu.append(z["cur_u"][a:e]); cv.append(z["cur_v"][a:e]); ru.append(r_u); rv.append(r_v); truth.append(ok); lens.append(n)
    a = off[0]
    for sl in (slice(a, a + 4), slice(a, a + 5)):                 # too few, exactly the minimal sample
        cu.append(z["cur_u"][sl]); cv.append(z["cur_v"][sl]); ru.append(z["ref_u"][sl]); rv.append(z["ref_v"][sl])
        truth.append(np.ones(sl.stop - sl.start, bool)); lens.append(sl.stop - sl.start)
    for arr, src in ((cu, "cur_u"), (cv, "cur_v"), (ru, "ref_u"), (rv, "ref_v")):      # rank deficient: one correspondence six times
        arr.append(np.repeat(z[src][a:a + 1], 6))
    truth.append(np.ones(6, bool)); lens.append(6)
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    cat = lambda x: np.concatenate(x).astype(np.float32)
    return dict(offsets=offsets, cur_u=cat(cu), cur_v=cat(cv), ref_u=cat(ru), ref_v=cat(rv), true_match=np.concatenate(truth),
                true_poses=z["true_poses"])


def main():
    d = build_inputs()
    off = d["offsets"]
    F = len(off) - 1
    E = np.zeros((F, 9)); mask = np.zeros(off[-1], np.uint8); cnt = np.zeros(F, np.int32); hyp = np.zeros(F, np.int32)
    for f in range(F):
        a, e = off[f], off[f + 1]
        cur = np.stack([d["cur_u"][a:e], d["cur_v"][a:e]], 1)
        ref = np.stack([d["ref_u"][a:e], d["ref_v"][a:e]], 1)
        Ef, m, c, h, _ = PL.find_essential_philox(cur, ref, *K, hypotheses=HYPOTHESES, threshold=THRESHOLD, seed=SEED, frame=f, seq=SEQ)
        E[f] = Ef.reshape(-1); mask[a:e] = m; cnt[f] = c; hyp[f] = h
        print(f, e - a, c, h, int(d["true_match"][a:e].sum()))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "essential.npz"), hypotheses=HYPOTHESES, threshold=THRESHOLD, seed=SEED, seq=SEQ,
                        E=E, mask=mask, n_inliers=cnt, best_hyp=hyp, **d)


if __name__ == "__main__":
    main()
