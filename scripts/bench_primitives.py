"""Device-timed micro-benchmarks of the stand-alone primitives (aux_kernels.cuh) at BASELINE-shaped sizes, each against the
roofline that bounds it.  Algorithmic bytes = the arrays the call must read and write once.  Prints one JSON document.
usage: python scripts/bench_primitives.py > gpurun_out/primitives_r01.json"""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvoscalerecovery_b200 import synth
from mvoscalerecovery_b200.batch import ScaleRecovery

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
PEAK = float(json.load(open(peaks))["hbm_gbs"]) if os.path.isfile(peaks) else 6650.0
eng = ScaleRecovery(absolute_reference=1.7)
dev = eng.device
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)           # > 126 MB L2


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    ms = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return float(np.median(ms))


out = {"peak_gbs": PEAK, "note": "median of 10 device-timed calls, L2 flushed between calls", "rows": []}
rng = np.random.default_rng(0)


def row(name, what, ms, alg_bytes, units, unit_name):
    out["rows"].append({"kernel": name, "workload": what, "ms": ms, "algorithmic_bytes": alg_bytes, "achieved_gbs": alg_bytes / ms / 1e6,
                        "frac_of_hbm_peak": alg_bytes / ms / 1e6 / PEAK, unit_name + "_per_s": units / ms * 1e3})


# ---- one KITTI-00-sized sequence: 4541 frames x ~2000 points, ~3960 triangles per frame
F, n = 4541, 2000
N, T = F * n, F * 3960
xyz = np.stack([rng.uniform(-8, 8, N), 1.7 + 0.01 * rng.standard_normal(N), rng.uniform(5, 40, N)], 1)
base = (np.arange(F) * n).repeat(3960)
tri = (rng.integers(0, n - 40, (T, 1)) + rng.integers(0, 40, (T, 3)) + base[:, None]).astype(np.int32)      # local triangles
d_tri, d_xyz = t(tri), t(xyz)
ms = timed(lambda: eng.triangle_planes(d_tri, d_xyz))
row("triangle_planes_kernel", "%d triangles over %d points (one 4541-frame sequence)" % (T, N), ms, 12 * T + 24 * N + 40 * T, T, "triangles")
d_v, d_d = t(np.ascontiguousarray(xyz[:, 1])), t(np.ascontiguousarray(xyz[:, 2]))
ms = timed(lambda: eng.triangle_votes(d_tri, d_v, d_d))
row("triangle_votes_kernel", "same mesh", ms, 12 * T + 16 * N + 8 * N, T, "triangles")

# ---- plane RANSAC: 4541 vertex lists of 6000 points, 100 hypotheses each (early stop as the reference)
M = 6000
pts = np.stack([rng.uniform(-8, 8, F * M), 1.7 + 0.003 * rng.standard_normal(F * M), rng.uniform(5, 40, F * M)], 1)
off = (np.arange(F + 1) * M).astype(np.int32)
d_off, d_pts = t(off), t(pts)
for stop in (True, False):
    ms = timed(lambda: eng.ransac_planes(d_off, d_pts, iterations=100, stop_at_goal=stop, seed=1), reps=5)
    row("ransac_planes_kernel", "%d lists x %d points, 100 hypotheses, %s" % (F, M, "early stop" if stop else "all hypotheses"), ms, 24 * F * M + 56 * F, F, "lists")

# ---- trajectories: the 11 KITTI-shaped sequences
lens = [4541, 1101, 4661, 801, 271, 2761, 1101, 1101, 4071, 1591, 1201]
so = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
FT = int(so[-1])
mot = np.tile(np.hstack([np.eye(3), [[0.0], [0.0], [1.0]]]).reshape(-1), (FT, 1)) + 1e-3 * rng.standard_normal((FT, 12))
d_so, d_mot, d_sc = t(so), t(mot), t(rng.uniform(0.5, 1.2, FT))
ms = timed(lambda: eng.integrate_paths(d_so, d_mot, d_sc))
row("integrate_paths_kernel", "11 sequences, %d frames" % FT, ms, 96 * FT + 8 * FT + 96 * (FT + 11), FT, "frames")

# ---- dense depth: one 1241x376 image, ~3960 triangles
from scipy.spatial import Delaunay
p2 = np.stack([rng.uniform(0, 1240, n), rng.uniform(186, 375, n)], 1)
tr = Delaunay(p2).simplices.astype(np.int32)
datas = np.hstack([np.tile([0.0, 1.0, 0.0], (tr.shape[0], 1)), np.full((tr.shape[0], 1), 1.7)])
d_tr, d_p2, d_da = t(tr), t(p2), t(datas)
ms = timed(lambda: eng.depth_from_mesh(1241, 376, 718.856, 718.856, 607.1928, 185.2157, d_tr, d_p2, d_da))
npx = 1241 * 376
row("raster_mesh_kernel + mesh_depth_kernel", "1241x376 pixels, %d triangles" % tr.shape[0], ms, 12 * npx + 44 * tr.shape[0] + 16 * n, npx, "pixels")

# ---- pose from E: 4541 frames x 2500 correspondences
b = synth.make_sequence(seed=1, n_frames=592, n_corr=2500)
E = np.zeros((b.n_frames, 9))
for f in range(b.n_frames):
    P = b.poses[f].reshape(3, 4); tt = P[:, 3] / np.linalg.norm(P[:, 3])
    E[f] = (np.array([[0, -tt[2], tt[1]], [tt[2], 0, -tt[0]], [-tt[1], tt[0], 0]]) @ P[:, :3]).reshape(-1)
dd = [t(x) for x in (b.offsets, b.cur_u, b.cur_v, b.ref_u, b.ref_v)]
d_E = t(E)
ms = timed(lambda: eng.recover_pose_frames(*dd, d_E))
row("recover_pose_kernel", "592 frames x 2500 correspondences (4 triangulations each)", ms, 16 * int(b.offsets[-1]) + (72 + 96 + 16) * b.n_frames, b.n_frames, "frames")

# ---- essential matrix by five-point RANSAC on the same 592 frames (FP64-pipe bound, not HBM: the fraction column only says so),
#      with OpenCV's own findEssentialMat on the host beside it (one core, a bounded sample of the same frames)
for hyps, conf in ((128, 0.0), (512, 0.0), (1000, 0.999)):
    ms = timed(lambda: eng.find_essential_frames(*dd, hypotheses=hyps, threshold=0.5, seed=1, confidence=conf), reps=5)
    row("find_essential_kernel", "592 frames x 2500 correspondences, %s %d hypotheses per frame" % ("up to" if conf else "exactly", hyps) +
        (" (confidence %.3f: the reference's call)" % conf if conf else ""), ms, 17 * int(b.offsets[-1]) + (72 + 8) * b.n_frames,
        b.n_frames, "frames")
    if not conf:
        out["rows"][-1]["hypotheses_per_s"] = hyps * b.n_frames / ms * 1e3
# ---- the whole chain from tracks alone: essential matrix -> pose -> fused stages 1-5 (the reference's call: maxIters 1000, prob 0.999)
maxf = int(np.diff(b.offsets).max())
ms = timed(lambda: eng.scale_frames_from_tracks(*dd, max_features=maxf, seed=1), reps=5)
row("find_essential_kernel + recover_pose_kernel + frame_kernel", "592 frames x 2500 correspondences, tracks -> raw scales", ms,
    16 * int(b.offsets[-1]) + 64 * b.n_frames, b.n_frames, "frames")
try:
    import time
    import cv2
    Kmat = np.array([[718.856, 0, 607.1928], [0, 718.856, 185.2157], [0, 0, 1.0]])
    t0 = time.perf_counter()
    for f in range(16):
        a, e = b.offsets[f], b.offsets[f + 1]
        cv2.findEssentialMat(np.stack([b.cur_u[a:e], b.cur_v[a:e]], 1), np.stack([b.ref_u[a:e], b.ref_v[a:e]], 1), cameraMatrix=Kmat,
                             method=cv2.RANSAC, prob=0.999, threshold=0.5)
    out["rows"][-1]["opencv_findEssentialMat_frames_per_s_one_core"] = 16 / (time.perf_counter() - t0)
except Exception as ex:                                             # the bench must not die on the host-side comparison
    out["rows"][-1]["opencv_findEssentialMat_frames_per_s_one_core"] = "unavailable: %r" % (ex,)
print(json.dumps(out, indent=1))
