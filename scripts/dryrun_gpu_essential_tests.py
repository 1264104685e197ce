"""Dry run of tests/test_gpu_zz_essential.py and tests/test_gpu_zzz_bucket.py WITHOUT a GPU -- a check of the test code and its tolerances, not of the device.
The seven GPU tests are called with a stand-in engine: find_essential_frames = the kernel SOURCE under the pthread emulation
(tests/host_sim/fp5_kernel_emu.cpp), the other entry points = the oracle (oracle.extras.recover_pose, oracle.pipeline).  Written
because the five-point kernel was finished after the round's GPU minutes were spent: a bug in an assertion or an index of the
test file would otherwise only surface on the B200.  Usage: python scripts/dryrun_gpu_essential_tests.py"""
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import numpy as np, torch                                            # noqa: E402
import test_five_point_host_sim as T                                 # noqa: E402
import test_gpu_zz_essential as G                                    # noqa: E402
from oracle import extras as X, pipeline as P                       # noqa: E402
from mvoscalerecovery_b200.batch import ScaleRecovery                # noqa: E402
from mvoscalerecovery_b200.compat import vo_geometry                 # noqa: E402

warnings.simplefilter("ignore")
emu = T.load_kernel_emulation(); K = T.K; f = torch.from_numpy


class StandInEngine:
    device = torch.device("cpu")
    scale_frames_from_tracks = ScaleRecovery.scale_frames_from_tracks

    def find_essential_frames(self, offsets, cur_u, cur_v, ref_u, ref_v, hypotheses=1000, threshold=0.5, seed=0, frame_index=None, seq_id=0,
                              confidence=0.999):
        off = np.ascontiguousarray(offsets.numpy()); F = len(off) - 1
        arr = [np.ascontiguousarray(x.numpy()) for x in (cur_u, cur_v, ref_u, ref_v)]
        E = np.zeros((F, 9)); mask = np.zeros(arr[0].size, np.uint8); cnt = np.zeros(F, np.int32); hyp = np.full(F, -1, np.int32)
        used = np.zeros(F, np.int32)
        fi = None if frame_index is None else np.ascontiguousarray(frame_index.numpy())
        emu.fp5_emu_find_essential(F, T._p(off), *(T._p(a) for a in arr), *K, hypotheses, threshold, confidence, seed,
                                   None if fi is None else T._p(fi), seq_id, T._p(E), T._p(mask), T._p(cnt), T._p(hyp), T._p(used), 3)
        return dict(essential=f(E), e_mask=f(mask), n_inliers=f(cnt), best_hyp=f(hyp), hyps_used=f(used))

    def _frames(self, offsets, cu, cv, ru, rv, e_mask):
        off = offsets.numpy()
        for fr in range(len(off) - 1):
            a, e = off[fr], off[fr + 1]
            cur = np.stack([cu.numpy()[a:e], cv.numpy()[a:e]], 1).astype(np.float64)
            ref = np.stack([ru.numpy()[a:e], rv.numpy()[a:e]], 1).astype(np.float64)
            yield fr, a, e, cur, ref, (np.ones(e - a, bool) if e_mask is None else e_mask.numpy()[a:e].astype(bool))

    def recover_pose_frames(self, offsets, cu, cv, ru, rv, essential, e_mask=None):
        F = offsets.numel() - 1
        poses = np.zeros((F, 12)); good = np.zeros((F, 4), np.int32)
        for fr, a, e, cur, ref, m in self._frames(offsets, cu, cv, ru, rv, e_mask):
            R, t, _, counts = X.recover_pose(essential.numpy()[fr].reshape(3, 3), cur[m], ref[m], *K)
            poses[fr] = np.hstack([R, np.asarray(t).reshape(3, 1)]).reshape(-1); good[fr] = counts
        return dict(poses=f(poses), n_good=f(good))

    def _triangulate(self, cur, ref, pose, m):
        Pm = pose.reshape(3, 4)
        Xw, ok = P.triangulate_dlt(cur, ref, Pm[:, :3], Pm[:, 3], *K)
        return Xw, ok & m

    def triangulate_frames(self, offsets, cu, cv, ru, rv, poses, e_mask=None):
        out = {k: np.zeros(cu.numel(), np.float32) for k in "xyzuv"}; n_out = np.zeros(offsets.numel() - 1, np.int32)
        for fr, a, e, cur, ref, m in self._frames(offsets, cu, cv, ru, rv, e_mask):
            Xw, ok = self._triangulate(cur, ref, poses.numpy()[fr], m)
            k = int(ok.sum()); n_out[fr] = k
            out["x"][a:a + k], out["y"][a:a + k], out["z"][a:a + k] = Xw[ok, 0], Xw[ok, 1], Xw[ok, 2]
        r = {k: f(v) for k, v in out.items()}; r["n_out"] = f(n_out)
        return r

    def pose_mask_frames(self, offsets, cu, cv, ru, rv, poses, e_mask=None):
        mask = np.zeros(cu.numel(), np.uint8)
        for fr, a, e, cur, ref, m in self._frames(offsets, cu, cv, ru, rv, e_mask):
            mask[a:e] = self._triangulate(cur, ref, poses.numpy()[fr], m)[1]
        return f(mask)

    def scale_frames_from_correspondences(self, offsets, cu, cv, ru, rv, poses, max_features, e_mask=None, frame_index0=0, seq_id=0, seed=0, stats=False):
        F = offsets.numel() - 1
        raw = np.full(F, np.nan); status = np.zeros(F, np.uint8)
        for fr, a, e, cur, ref, m in self._frames(offsets, cu, cv, ru, rv, e_mask):
            Xw, ok = self._triangulate(cur, ref, poses.numpy()[fr], m)
            f3 = Xw[ok].astype(np.float32).astype(np.float64); f2 = P.reproject(f3, K[0], K[2], K[3]).astype(np.float32).astype(np.float64)
            rec = P.frame_raw_scale(f3, f2, seed, frame_index0 + fr, seq_id, absolute_reference=1.7)
            raw[fr] = rec["raw_scale"]; status[fr] = 1 if rec["updated"] else 0
        return dict(raw_scale=f(raw), status=f(status), n_features=f(np.zeros(F, np.int32)), stats=None)


torch.cuda.synchronize = lambda *a, **k: None
eng = StandInEngine()
vo_geometry._engine = lambda Km: eng
z = np.load(os.path.join(ROOT, "tests", "golden", "essential.npz"))
d = {k: G._t(eng, z[k]) for k in ("offsets", "cur_u", "cur_v", "ref_u", "ref_v")}
out = eng.find_essential_frames(d["offsets"], d["cur_u"], d["cur_v"], d["ref_u"], d["ref_v"], hypotheses=int(z["hypotheses"]),
                                threshold=float(z["threshold"]), seed=int(z["seed"]), seq_id=int(z["seq"]))
run = (d, {k: v.cpu().numpy() for k, v in out.items()})
G.test_self_consistency_and_edge_frames(z, run)
G.test_against_the_oracle_golden_and_the_truth(z, run)
G.test_against_the_host_build_of_the_kernel_numerics(z, run)
G.test_pose_from_the_gpu_essential_matrix(eng, z, run)
G.test_hypothesis_count_and_frame_index(eng, z, run)
G.test_tracks_to_scales_without_poses(eng)
G.test_process_tracks_drop_in(eng, z)
print("all 7 tests of tests/test_gpu_zz_essential.py pass against the CPU stand-ins")

# ---- the GPU tests of the bucketing kernel (tests/test_gpu_zzz_bucket.py) against the emulation of bucket_kernel's source
import test_bucket as TB                                             # noqa: E402
import test_gpu_zzz_bucket as GB                                     # noqa: E402
from mvoscalerecovery_b200.compat import _gpu as compat_gpu          # noqa: E402

bemu = TB.load_bucket_emulation()


def bucket_frames(self, offsets, u, v, bucket_size=30, density=2, seed=0, frame_index=None, seq_id=0):
    off = np.ascontiguousarray(offsets.numpy()); F = len(off) - 1
    uu, vv = np.ascontiguousarray(u.numpy()), np.ascontiguousarray(v.numpy())
    fi = None if frame_index is None else np.ascontiguousarray(frame_index.numpy())
    index = np.full(uu.size, -1, np.int32); n_out = np.zeros(F, np.int32); status = np.zeros(F, np.uint8)
    bemu.bucket_emu(F, TB._p(off), TB._p(uu), TB._p(vv), bucket_size, density, seed, TB._p(fi), seq_id, TB._p(index), TB._p(n_out), TB._p(status), 2)
    return dict(index=f(index), n_out=f(n_out), status=f(status))


StandInEngine.bucket_frames = bucket_frames
compat_gpu.engine = lambda *a, **k: eng
gb = np.load(os.path.join(ROOT, "tests", "golden", "bucket.npz"))
GB.test_bucket_frames_on_the_gpu(eng, gb)
GB.test_compat_bucket_returns_what_the_reference_returns(eng, gb)
print("both tests of tests/test_gpu_zzz_bucket.py pass against the CPU stand-in")
