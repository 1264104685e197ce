"""The slowest frames of the bench sequence: per-phase cycles and star-path counters (profiling aid, round 2)."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvoscalerecovery_b200 import synth, _native as N
from mvoscalerecovery_b200.batch import ScaleRecovery, stats_to_numpy
NAMES = {0: "load", 1: "grid1", 2: "stars1", 3: "s1pair", 6: "compact+grid2", 7: "stars2", 8: "s2pair", 9: "nwrap", 10: "planes", 11: "median", 12: "list", 13: "ransac"}
n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 4541
b = synth.make_sequence(seed=20261017, n_frames=n_frames, n_corr=2500, outlier_frac=0.10)
eng = ScaleRecovery(absolute_reference=1.7)
dev = eng.device
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
d = [t(x) for x in (b.offsets, b.cur_u, b.cur_v, b.ref_u, b.ref_v, b.poses)]
maxf = int(np.max(np.diff(b.offsets)))
ph = torch.zeros(n_frames * 16, dtype=torch.int64, device=dev)
for it in range(3):
    if it == 2:
        N.check(eng.lib.mvosr_set_phase_timing(eng._h, C.c_void_p(ph.data_ptr())))
    out = eng.scale_frames_from_correspondences(*d, max_features=maxf, seed=1, stats=True)
    torch.cuda.synchronize()
N.check(eng.lib.mvosr_set_phase_timing(eng._h, None))
p = ph.cpu().numpy().reshape(n_frames, 16).astype(np.float64)
st = stats_to_numpy(out["stats"])
tot = p[:, [0, 1, 2, 6, 7, 10, 11, 12, 13]].sum(1)
print("frames over 1.6M: %d, over 2M: %d, over 3M: %d; mean %.0f" % ((tot > 1.6e6).sum(), (tot > 2e6).sum(), (tot > 3e6).sum(), tot.mean()))
for f in np.argsort(-tot)[:12]:
    print("f%-5d tot %.2fM " % (f, tot[f] / 1e6) + " ".join("%s %.0fk" % (NAMES[k], p[f, k] / 1e3) if k != 9 else "nwrap %d" % p[f, k] for k in NAMES) +
          " | n_roi %d kept %d deferred %d fb %d exact %d hyps %d" % (st["n_roi"][f], st["n_kept"][f], st["n_deferred"][f] & 0xFFFF, st["n_deferred"][f] >> 16, st["n_exact"][f], st["hyps_used"][f]))
