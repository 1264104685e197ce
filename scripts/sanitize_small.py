"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck): fused path, staged path with debug buffers,
large-frame mode, Delaunay-only, filter, stand-alone primitives."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvoscalerecovery_b200 import synth
from mvoscalerecovery_b200.batch import ScaleRecovery

eng = ScaleRecovery(absolute_reference=1.7)
dev = eng.device
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
for n_corr, nf in ((900, 6), (5200, 2)):                       # shared-memory mode, large-frame mode
    b = synth.make_sequence(seed=5, n_frames=nf, n_corr=n_corr, outlier_frac=0.2)
    maxf = int(np.max(np.diff(b.offsets)))
    r = eng.scale_frames_from_correspondences(t(b.offsets), t(b.cur_u), t(b.cur_v), t(b.ref_u), t(b.ref_v), t(b.poses), max_features=maxf, seed=3, stats=True)
    s1 = eng.triangulate_frames(t(b.offsets), t(b.cur_u), t(b.cur_v), t(b.ref_u), t(b.ref_v), t(b.poses))
    r2 = eng.scale_frames(t(b.offsets), s1["x"], s1["y"], s1["z"], s1["u"], s1["v"], maxf, counts=s1["n_out"], seed=3, debug=True)
    dt = eng.delaunay_frames(t(b.offsets), s1["u"], s1["v"], maxf)
    seq = torch.tensor([0, nf], dtype=torch.int32, device=dev)
    f = eng.filter_sequences(seq, r["raw_scale"], r["status"], t(b.move_flags), r["n_features"])
    p = eng.integrate_paths(seq, t(b.poses), f["scale"])
    torch.cuda.synchronize()
    assert np.array_equal(r["raw_scale"].cpu().numpy(), r2["raw_scale"].cpu().numpy(), equal_nan=True)
    print("ok", n_corr, r["raw_scale"].cpu().numpy()[:3], b.true_scale[:3])
rng = np.random.default_rng(1)
pts = np.stack([rng.uniform(-8, 8, 500), 1.7 + 0.003 * rng.standard_normal(500), rng.uniform(5, 40, 500)], 1)
out = eng.ransac_planes(t(np.array([0, 500], np.int32)), t(pts), iterations=64)
torch.cuda.synchronize()
print("ransac ic", int(out["ic"][0]))
# five-point RANSAC + pose + fused scale recovery from tracks alone (frames of several tiles, a frame below five correspondences)
b = synth.make_sequence(seed=9, n_frames=4, n_corr=700, outlier_frac=0.1)
off = np.concatenate([b.offsets, [b.offsets[-1] + 3]]).astype(np.int32)            # a fifth frame of three correspondences
pad = lambda a: np.concatenate([a, a[:3]])
r = eng.scale_frames_from_tracks(t(off), t(pad(b.cur_u)), t(pad(b.cur_v)), t(pad(b.ref_u)), t(pad(b.ref_v)), max_features=int(np.diff(off).max()),
                                 hypotheses=160, seed=3, confidence=0.0)
torch.cuda.synchronize()
print("essential inliers", r["n_inliers"].cpu().numpy(), "scales", r["raw_scale"].cpu().numpy()[:4], "true", b.true_scale)
m = eng.pose_mask_frames(t(off), t(pad(b.cur_u)), t(pad(b.cur_v)), t(pad(b.ref_u)), t(pad(b.ref_v)), r["poses"], e_mask=r["e_mask"])
torch.cuda.synchronize()
print("pose mask", int(m.sum()), "of", m.numel())
bk = eng.bucket_frames(t(off), t(pad(b.cur_u)), t(pad(b.cur_v)), seed=3)
torch.cuda.synchronize()
print("bucketing kept", bk["n_out"].cpu().numpy())
