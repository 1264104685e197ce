"""BASELINE configs[4]: RANSAC sweep -- 256..4096 Philox hypotheses/frame x 0..40 % outlier correspondences.
Reports frames/s of the fused kernel (CUDA events) and the per-frame scale error against the synthetic truth, with the
reference's early stop (first hypothesis above 0.8 N inliers) and with all hypotheses evaluated.
usage: python scripts/ransac_sweep.py [n_frames] > gpurun_out/ransac_sweep.json"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvoscalerecovery_b200 import synth                         # noqa: E402
from mvoscalerecovery_b200.batch import ScaleRecovery            # noqa: E402


def main():
    n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 592
    rows = []
    for outl in (0.0, 0.1, 0.2, 0.3, 0.4):
        b = synth.make_sequence(seed=4242, n_frames=n_frames, n_corr=2500, outlier_frac=outl)
        for H in (100, 256, 512, 1024, 2048, 4096):
            for stop in (1, 0):
                eng = ScaleRecovery(absolute_reference=1.7, ransac_iterations=H, ransac_stop_at_goal=stop)
                dev = eng.device
                t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
                d = [t(x) for x in (b.offsets, b.cur_u, b.cur_v, b.ref_u, b.ref_v, b.poses)]
                maxf = int(np.max(np.diff(b.offsets)))
                ms = []
                for it in range(4):
                    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                    e0.record()
                    out = eng.scale_frames_from_correspondences(*d, max_features=maxf, seed=1)
                    e1.record(); torch.cuda.synchronize()
                    if it:
                        ms.append(e0.elapsed_time(e1))
                raw = out["raw_scale"].cpu().numpy(); st = out["status"].cpu().numpy()
                okm = (st & 1) != 0
                err = np.abs(raw[okm] - b.true_scale[okm]) / b.true_scale[okm]
                rows.append(dict(outlier_frac=outl, hypotheses=H, stop_at_goal=stop, fps=n_frames / (min(ms) * 1e-3),
                                 updated_frac=float(okm.mean()), err_median=float(np.median(err)), err_p95=float(np.percentile(err, 95))))
                eng.close()
    print(json.dumps(dict(n_frames=n_frames, correspondences_per_frame=2500, rows=rows)))


if __name__ == "__main__":
    main()
