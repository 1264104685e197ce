"""Per-phase SM-cycle breakdown of the fused frame kernel (uses mvosr_set_phase_timing). Profiling aid."""
import ctypes as C
import os
import sys
import json

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvoscalerecovery_b200 import synth, _native as N            # noqa: E402
from mvoscalerecovery_b200.batch import ScaleRecovery, stats_to_numpy            # noqa: E402

NAMES = {0: "load+stage1+roi", 1: "grid1", 2: "stars1", 3: "stars1_pair", 6: "keep+compact+grid2", 7: "stars2", 8: "stars2_pair", 9: "n_to_wrap_path", 10: "planes", 11: "median", 12: "valid_list", 13: "ransac"}


def main():
    n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 592
    n_corr = int(sys.argv[2]) if len(sys.argv) > 2 else 2500
    b = synth.make_sequence(seed=20261017, n_frames=n_frames, n_corr=n_corr, outlier_frac=0.10, **({'density': os.environ['MVOSR_DATA']} if os.environ.get('MVOSR_DATA') else {}))
    kw = {}
    if os.environ.get('MVOSR_DENSITY'):
        kw['grid_density'] = float(os.environ['MVOSR_DENSITY'])
    if os.environ.get('MVOSR_WFAC'):
        kw['window_factor'] = float(os.environ['MVOSR_WFAC'])
    eng = ScaleRecovery(absolute_reference=1.7, **kw)
    dev = eng.device
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d = [t(x) for x in (b.offsets, b.cur_u, b.cur_v, b.ref_u, b.ref_v, b.poses)]
    maxf = int(np.max(np.diff(b.offsets)))
    ph = torch.zeros(n_frames * 16, dtype=torch.int64, device=dev)
    for it in range(3):
        if it == 2:
            N.check(eng.lib.mvosr_set_phase_timing(eng._h, C.c_void_p(ph.data_ptr())))
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        out = eng.scale_frames_from_correspondences(*d, max_features=maxf, seed=1, stats=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    N.check(eng.lib.mvosr_set_phase_timing(eng._h, None))
    p = ph.cpu().numpy().reshape(n_frames, 16).astype(np.float64)
    st = stats_to_numpy(out["stats"])
    tot = p.sum(1).mean()
    res = {"frames": n_frames, "kernel_ms": ms, "fps": n_frames / ms * 1e3, "cycles_per_frame_total": tot,
           "phases": {NAMES[k]: {"cycles": float(p[:, k].mean()), "share": float(p[:, k].mean() / tot)} for k in sorted(NAMES)},
           "n_roi": float(st["n_roi"].mean()), "n_kept": float(st["n_kept"].mean()), "n_tri": float(st["n_tri"].mean()),
           "n_deferred": float((st["n_deferred"] & 0xFFFF).mean()), "n_fallback": float((st["n_deferred"] >> 16).mean()), "hyps_used": float(st["hyps_used"].mean()), "n_exact": float(st["n_exact"].mean())}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
