"""Raw per-frame means of the 16 phase/counter slots (for -DMVOSR_STAR_COUNTERS / -DMVOSR_WRAP_COUNTERS builds)."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvoscalerecovery_b200 import synth, _native as N
from mvoscalerecovery_b200.batch import ScaleRecovery
n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 592
b = synth.make_sequence(seed=20261017, n_frames=n_frames, n_corr=2500, outlier_frac=0.10)
eng = ScaleRecovery(absolute_reference=1.7); dev = eng.device
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
d = [t(x) for x in (b.offsets, b.cur_u, b.cur_v, b.ref_u, b.ref_v, b.poses)]
ph = torch.zeros(n_frames * 16, dtype=torch.int64, device=dev)
N.check(eng.lib.mvosr_set_phase_timing(eng._h, C.c_void_p(ph.data_ptr())))
eng.scale_frames_from_correspondences(*d, max_features=int(np.max(np.diff(b.offsets))), seed=1, stats=True)
torch.cuda.synchronize()
p = ph.cpu().numpy().reshape(n_frames, 16).astype(np.float64).mean(0)
# cnt[0..7] -> slots 4,5,11,12,14,15,1,10
print("cnt0..7:", " ".join("%.1f" % p[k] for k in (4, 5, 11, 12, 14, 15, 1, 10)), " to_wrap %.1f" % p[9])
