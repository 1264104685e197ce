"""Raw per-frame means of the 16 phase/counter slots (for -DMVOSR_STAR_COUNTERS / -DMVOSR_WRAP_COUNTERS builds).
MVOSR_DATA=uniform|ground|clustered selects the feature density; MVOSR_WRAP=1 prints the wrap-path summary of a
-DMVOSR_WRAP_COUNTERS build, otherwise the pair path's reasons for giving a star up (-DMVOSR_STAR_COUNTERS)."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvoscalerecovery_b200 import synth, _native as N
from mvoscalerecovery_b200.batch import ScaleRecovery
n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 592
kw = {'density': os.environ['MVOSR_DATA']} if os.environ.get('MVOSR_DATA') else {}
b = synth.make_sequence(seed=20261017, n_frames=n_frames, n_corr=2500, outlier_frac=0.10, **kw)
eng = ScaleRecovery(absolute_reference=1.7); dev = eng.device
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
d = [t(x) for x in (b.offsets, b.cur_u, b.cur_v, b.ref_u, b.ref_v, b.poses)]
ph = torch.zeros(n_frames * 16, dtype=torch.int64, device=dev)
N.check(eng.lib.mvosr_set_phase_timing(eng._h, C.c_void_p(ph.data_ptr())))
eng.scale_frames_from_correspondences(*d, max_features=int(np.max(np.diff(b.offsets))), seed=1, stats=True)
torch.cuda.synchronize()
p = ph.cpu().numpy().reshape(n_frames, 16).astype(np.float64).mean(0)
# cnt[0..7] -> slots 4,5,11,12,14,15,1,10
c = [p[k] for k in (4, 5, 11, 12, 14, 15, 1, 10)]
print("cnt0..7:", " ".join("%.1f" % v for v in c), " to_wrap %.1f" % p[9])
if os.environ.get('MVOSR_WRAP'):       # steps, stream calls, stars that streamed, open stars, cycles, stars, cycles of open stars, cycles in w_stream
    print("wrap stars %.1f (open %.1f, streamed %.1f)  steps/star %.2f  stream calls/star %.2f  warp-cycles/star %.0f (open stars %.0f, closed %.0f)  in w_stream %.0f%%  sum warp-cycles/28 %.0f" % (
        c[5], c[3], c[2], c[0] / c[5], c[1] / c[5], c[4] / c[5], c[6] / max(c[3], 1), (c[4] - c[6]) / max(c[5] - c[3], 1), 100 * c[7] / c[4], c[4] / 28))
else:                                  # pair path, both passes: why a star was given up
    print("pair path gives up per frame: crowded block (> 64 candidates) %.1f, nearest point beyond the block margin %.1f, nearest not unique / not closed %.1f, "
          "side not certain %.1f, nothing on the left (hull / far neighbour) %.1f, interval clash %.1f, cap leaves the block %.1f, more than 16 neighbours %.1f" % tuple(c))
