"""Why does the processing order of the persistent CTAs change the kernel time?  Per-phase SM cycles of the fused frame kernel
(mvosr_set_phase_timing) for several orders of the same 4541 frames.  Profiling aid (round 2)."""
import ctypes as C, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvoscalerecovery_b200 import synth, _native as N
from mvoscalerecovery_b200.batch import ScaleRecovery
NAMES = {0: "load", 1: "grid1", 2: "stars1", 6: "compact+grid2", 7: "stars2", 10: "planes", 11: "median", 12: "list", 13: "ransac"}
n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 4541
b = synth.make_sequence(seed=20261017, n_frames=n_frames, n_corr=2500, outlier_frac=0.10)
eng = ScaleRecovery(absolute_reference=1.7)
dev = eng.device
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
d = [t(x) for x in (b.offsets, b.cur_u, b.cur_v, b.ref_u, b.ref_v, b.poses)]
maxf = int(np.max(np.diff(b.offsets)))
sizes = np.diff(b.offsets)
rec = torch.zeros(n_frames, 16, dtype=torch.uint8, device=dev)
srt = np.argsort(-sizes, kind="stable")
orders = {"natural": None, "lpt": srt, "random": np.random.default_rng(1).permutation(n_frames), "strided": np.concatenate([srt[k::31] for k in range(31)]),
          "reverse": np.arange(n_frames)[::-1].copy(), "blocked-random": np.concatenate([np.random.default_rng(2).permutation(np.arange(a, min(a + 148, n_frames))) for a in range(0, n_frames, 148)])}
ph = torch.zeros(n_frames * 16, dtype=torch.int64, device=dev)
for name, o in orders.items():
    od = None if o is None else t(o.astype(np.int32))
    ms = []
    for it in range(6):
        if it == 5:
            N.check(eng.lib.mvosr_set_phase_timing(eng._h, C.c_void_p(ph.data_ptr())))
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.scale_shard_from_correspondences(*d, maxf, rec, order=od, seed=1)
        e1.record(); torch.cuda.synchronize()
        if 2 <= it < 5:
            ms.append(e0.elapsed_time(e1))
    N.check(eng.lib.mvosr_set_phase_timing(eng._h, None))
    p = ph.cpu().numpy().reshape(n_frames, 16).astype(np.float64)
    tot = p[:, list(NAMES)].sum(1)
    print("%-15s kernel %.3f ms  cycles/frame mean %.0f p50 %.0f p99 %.0f max %.0f | " % (name, np.mean(ms), tot.mean(), np.median(tot), np.percentile(tot, 99), tot.max())
          + " ".join("%s %.0f" % (NAMES[k], p[:, k].mean()) for k in NAMES))
