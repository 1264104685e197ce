"""The numerics of the five-point kernel without a GPU: csrc/five_point.cuh is __host__ __device__, tests/host_sim/fp5_host.cpp
compiles those very functions with g++ and this file checks them against the oracle (oracle/five_point.py = LAPACK version,
oracle/five_point_plan.py = the independent Python restatement on the shared Philox stream) and against the golden the GPU
test uses (tests/golden/essential.npz).  The host build is test infrastructure: nothing in the product loads it."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import five_point as FP
from oracle import five_point_plan as PL
from oracle import philox as PH
from mvoscalerecovery_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIM = os.path.join(ROOT, "tests", "host_sim")
K = (718.856, 718.856, 607.1928, 185.2157)
vp = C.c_void_p


def load_host_sim():
    """Builds (g++) and loads the host compilation of csrc/five_point.cuh.  Also used by tests/test_gpu_zz_essential.py."""
    so, src = os.path.join(SIM, "libfp5_host.so"), os.path.join(SIM, "fp5_host.cpp")
    deps = [src] + [os.path.join(ROOT, "mvoscalerecovery_b200", "csrc", f) for f in ("five_point.cuh", "five_point_tables.h")]
    if not os.path.isfile(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, src])
    L = C.CDLL(so)
    L.fp5_host_solve.restype = C.c_int
    L.fp5_host_solve.argtypes = [vp, vp, vp]
    L.fp5_host_sample5.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, vp]
    L.fp5_host_philox.argtypes = [C.c_uint32] * 6 + [vp]
    L.fp5_host_ransac_frame.argtypes = [C.c_int32, vp, vp, vp, vp] + [C.c_double] * 4 + [C.c_int32, C.c_double, C.c_double, C.c_uint64, C.c_uint32, C.c_uint32,
                                                                                       vp, vp, vp, vp, vp]
    return L


@pytest.fixture(scope="module")
def sim():
    return load_host_sim()


def load_kernel_emulation():
    """Builds (g++ -pthread) and loads tests/host_sim/fp5_kernel_emu.cpp: the SOURCE of find_essential_kernel run on the host, one
    OS thread per CUDA thread."""
    so, src = os.path.join(SIM, "libfp5_kernel_emu.so"), os.path.join(SIM, "fp5_kernel_emu.cpp")
    deps = [src] + [os.path.join(ROOT, "mvoscalerecovery_b200", "csrc", f) for f in ("five_point.cuh", "five_point_kernel.cuh", "five_point_tables.h")]
    if not os.path.isfile(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-pthread", "-Wno-unknown-pragmas", "-o", so, src])
    L = C.CDLL(so)
    L.fp5_emu_find_essential.argtypes = [C.c_int32, vp, vp, vp, vp, vp] + [C.c_double] * 4 + [C.c_int32, C.c_double, C.c_double, C.c_uint64, vp, C.c_int32,
                                                                                         vp, vp, vp, vp, vp, C.c_int32]
    return L


def _p(a):
    return a.ctypes.data_as(vp)


def _solve(L, x1, x2):
    x1 = np.ascontiguousarray(x1, dtype=np.float64); x2 = np.ascontiguousarray(x2, dtype=np.float64)
    out = np.zeros((10, 9))
    n = L.fp5_host_solve(_p(x1), _p(x2), _p(out))
    return [out[k].reshape(3, 3).copy() for k in range(n)]


def _dist(E, S):
    return min([min(np.linalg.norm(E - F), np.linalg.norm(E + F)) for F in S] or [9.0])


def test_philox_kats_and_sample_stream(sim):
    out = np.zeros(4, np.uint32)
    for ctr, key, want in (((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),      # Random123 known answers
                           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
                           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))):
        sim.fp5_host_philox(*ctr, *key, _p(out))
        assert tuple(int(v) for v in out) == want == PH.philox4x32_10(ctr, key)
    idx = np.zeros(5, np.int32)
    for seed, hyp, frame, seq, n in ((0, 0, 0, 0, 5), (7, 3, 11, 2, 6), (2**63 + 12345, 4095, 4540, 10, 2000), (99, 17, 5, 0, 20000), (1, 2, 3, 4, 7)):
        sim.fp5_host_sample5(seed, hyp, frame, seq, n, _p(idx))
        assert list(idx) == PL.sample5_positions(seed, hyp, frame, seq, n)
        assert len(set(idx)) == 5 and 0 <= idx.min() and idx.max() < n
    seen = set()
    for hyp in range(400):                                         # n = 5: every draw is a permutation of range(5); all positions get used
        sim.fp5_host_sample5(5, hyp, 0, 0, 5, _p(idx))
        assert sorted(idx) == [0, 1, 2, 3, 4]
        seen.add(tuple(idx))
    assert len(seen) > 100


def _problem(rng, noisy):
    R = synth._rodrigues(*rng.uniform(-0.2, 0.2, 3))
    t = rng.standard_normal(3); t /= np.linalg.norm(t)
    P = np.stack([rng.uniform(-2, 2, 5), rng.uniform(-1, 1, 5), rng.uniform(4, 20, 5)], 1)
    x1 = P[:, :2] / P[:, 2:] + 1e-3 * rng.standard_normal((5, 2)) * noisy
    P2 = P @ R.T + t
    x2 = P2[:, :2] / P2[:, 2:]
    Et = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]]) @ R
    return x1, x2, Et / np.linalg.norm(Et)


def test_minimal_solver_against_lapack_and_plan(sim):
    rng = np.random.default_rng(3)
    lap = found = true_found = n_true = same_as_plan = 0
    trials = 60
    for trial in range(trials):
        x1, x2, Et = _problem(rng, trial % 2)
        a, h = FP.five_point(x1, x2), _solve(sim, x1, x2)
        assert len(h) <= 10
        for E in a:
            lap += 1
            found += _dist(E, h) < 1e-6
        for F in h:                                               # nothing spurious, every solution is an essential matrix through the five points
            assert _dist(F, a) < 1e-3
            assert abs(np.linalg.norm(F) - 1) < 1e-12
            s = np.linalg.svd(F, compute_uv=False)
            assert abs(s[0] - s[1]) < 1e-5 and s[2] < 1e-5
            x1h, x2h = np.hstack([x1, np.ones((5, 1))]), np.hstack([x2, np.ones((5, 1))])
            assert np.abs(np.sum(x2h * (x1h @ F.T), 1)).max() < 1e-6
        if trial % 2 == 0:
            n_true += 1
            true_found += _dist(Et, h) < 1e-7
        b = PL.five_point_device_style(x1, x2)
        same_as_plan += len(b) == len(h) and all(np.abs(F - G).max() < 1e-6 for F, G in zip(h, b))
    assert found >= 0.98 * lap and true_found >= 0.96 * n_true
    assert same_as_plan >= 0.95 * trials        # same algorithm family, independent code: a near-double root may be a real pair on
                                                # one side and a complex pair on the other


def test_degenerate_samples_give_no_solution_and_no_nan(sim):
    x = np.tile(np.array([[0.1, 0.2]]), (5, 1))
    assert _solve(sim, x, x) == []                                # one correspondence five times: rank deficient
    z = np.zeros((5, 2))
    assert _solve(sim, z, z) == []
    rng = np.random.default_rng(0)
    a = rng.uniform(-0.5, 0.5, (5, 2))
    for F in _solve(sim, a, a):                                   # pure rotation by identity: anything returned is finite and essential
        assert np.isfinite(F).all()
    bad = a.copy(); bad[2, 0] = np.nan
    for F in _solve(sim, bad, a):
        assert np.isfinite(F).all()


def _ransac(L, cu, cv, ru, rv, hyps, thr, seed, frame, seq, confidence=0.0, with_used=False):
    n = cu.size
    E = np.zeros(9); mask = np.zeros(max(n, 1), np.uint8); cnt = C.c_int32(0); hyp = C.c_int32(0); used = C.c_int32(0)
    cu, cv, ru, rv = (np.ascontiguousarray(x, dtype=np.float32) for x in (cu, cv, ru, rv))
    L.fp5_host_ransac_frame(n, _p(cu), _p(cv), _p(ru), _p(rv), *K, hyps, thr, confidence, seed, frame, seq, _p(E), _p(mask), C.byref(cnt), C.byref(hyp),
                            C.byref(used))
    out = (E.reshape(3, 3), mask[:n].astype(bool), cnt.value, hyp.value)
    return out + (used.value,) if with_used else out


def test_selection_rule_matches_the_oracle_exactly_given_the_same_candidates(sim):
    """find_essential_philox with the host build as its minimal solver: sampling, scoring, tie-breaking and the mask must agree
    bit for bit with the sequential replay of the kernel's rule."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "essential.npz"))
    off = z["offsets"]
    for f in (1, 4, 8, 9, 10):
        a, e = off[f], off[f + 1]
        cu, cv, ru, rv = (z[k][a:e] for k in ("cur_u", "cur_v", "ref_u", "ref_v"))
        E, mask, cnt, hyp = _ransac(sim, cu, cv, ru, rv, 24, 0.5, 77, f, 1)
        Eo, mo, co, ho, _ = PL.find_essential_philox(np.stack([cu, cv], 1), np.stack([ru, rv], 1), *K, hypotheses=24, threshold=0.5, seed=77, frame=f, seq=1,
                                                  solver=lambda x1, x2: _solve(sim, x1, x2))
        assert (cnt, hyp) == (co, ho) and np.array_equal(mask, mo) and np.array_equal(E, Eo)
        assert cnt == int(mask.sum())


def test_host_replay_against_the_golden_of_the_python_oracle(sim):
    """The golden of the GPU test (independent Python solver): same winner wherever both solvers found the same candidates; in
    every frame an inlier set of the same size within a point or two, and the true matches recovered."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "essential.npz"))
    off = z["offsets"]
    H, thr, seed, seq = int(z["hypotheses"]), float(z["threshold"]), int(z["seed"]), int(z["seq"])
    same = 0
    F = len(off) - 1
    for f in range(F):
        a, e = off[f], off[f + 1]
        E, mask, cnt, hyp = _ransac(sim, *(z[k][a:e] for k in ("cur_u", "cur_v", "ref_u", "ref_v")), H, thr, seed, f, seq)
        if e - a < 8:
            assert (cnt, hyp) == (int(z["n_inliers"][f]), int(z["best_hyp"][f]))
            continue
        assert abs(cnt - int(z["n_inliers"][f])) <= 2
        truth = z["true_match"][a:e]
        assert (mask & truth).sum() >= 0.97 * truth.sum() and (mask & ~truth).sum() <= 3
        if hyp == int(z["best_hyp"][f]):
            same += 1
            assert _dist(E, [z["E"][f].reshape(3, 3)]) < 1e-6 and np.array_equal(mask, z["mask"][a:e].astype(bool))
    assert same >= 6


def test_kernel_source_on_the_host_emulation(sim):
    """find_essential_kernel itself (five_point_kernel.cuh compiled for the host with a pthread mapping of the CUDA vocabulary):
    tile staging, barriers, the warp-shuffle key reduction, winner hand-over between rounds, frame loop over a small grid and the
    mask pass give, bit for bit, what the sequential replay gives -- for several grid sizes, partial rounds, frames of several
    tiles, a frame_index map and the edge frames.  (scripts/tsan_five_point_kernel.sh runs the same build under ThreadSanitizer.)"""
    emu = load_kernel_emulation()
    z = np.load(os.path.join(ROOT, "tests", "golden", "essential.npz"))
    big = synth.make_sequence(seed=5, n_frames=2, n_corr=1300, outlier_frac=0.1)          # 3 tiles per frame
    lens = list(np.diff(z["offsets"])) + list(np.diff(big.offsets))
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    arr = {k: np.ascontiguousarray(np.concatenate([z[k], getattr(big, k)]).astype(np.float32)) for k in ("cur_u", "cur_v", "ref_u", "ref_v")}
    rng = np.random.default_rng(1)
    bad = rng.permutation(arr["ref_u"].size)[: arr["ref_u"].size // 6]
    arr["ref_u"][bad] = rng.uniform(0, 1241, bad.size).astype(np.float32)
    F = len(off) - 1
    fidx = np.ascontiguousarray((np.arange(F) * 3 + 1).astype(np.int32))
    seed, seq = 2**40 + 17, 6
    for H, grid, use_index, conf in ((48, 4, False, 0.0), (150, 1, True, 0.0), (128, F, True, 0.999), (400, 3, False, 0.999), (300, 2, True, 0.5)):
        E = np.full((F, 9), 7.0); mask = np.full(off[-1], 9, np.uint8); cnt = np.full(F, -5, np.int32); hyp = np.full(F, -5, np.int32)
        used = np.full(F, -5, np.int32)
        rc = emu.fp5_emu_find_essential(F, _p(off), *(_p(arr[k]) for k in ("cur_u", "cur_v", "ref_u", "ref_v")), *K, H, 0.5, conf, seed,
                                        _p(fidx) if use_index else None, seq, _p(E), _p(mask), _p(cnt), _p(hyp), _p(used), grid)
        assert rc == 0
        for f in range(F):
            a, e = off[f], off[f + 1]
            Er, mr, cr, hr, ur = _ransac(sim, *(arr[k][a:e] for k in ("cur_u", "cur_v", "ref_u", "ref_v")), H, 0.5, seed,
                                         int(fidx[f]) if use_index else f, seq, confidence=conf, with_used=True)
            assert (cr, hr, ur) == (cnt[f], hyp[f], used[f]), (H, grid, f)
            assert np.array_equal(Er.reshape(-1), E[f]) and np.array_equal(mr, mask[a:e].astype(bool))
        assert cnt[-1] > 800 and cnt[-2] > 800                     # the large frames found their model
        if H == 400:                                               # ~80 % inliers: the adaptive rule stops after the first round
            assert (used[:8] == 128).all() and used[8] == 0 and used[9] == 128 and used[10] == 400


def test_adaptive_stopping_rule_against_the_oracle(sim):
    """OpenCV's adaptive hypothesis count (prob = 0.999), evaluated per round of 128: the replay of the kernel's rule and the
    oracle's stop at the same hypothesis, on frames from 90 % down to 35 % inliers, and the count is the formula's."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "pose.npz"))
    a, e = z["offsets"][2], z["offsets"][3]
    rng = np.random.default_rng(11)
    for frac_bad, want_rounds in ((0.0, 1), (0.45, 2), (0.65, None)):
        cu, cv, ru, rv = (z[k][a:e].copy() for k in ("cur_u", "cur_v", "ref_u", "ref_v"))
        bad = rng.permutation(e - a)[: int(frac_bad * (e - a))]
        ru[bad] = rng.uniform(0, 1241, bad.size).astype(np.float32); rv[bad] = rng.uniform(0, 376, bad.size).astype(np.float32)
        E, mask, cnt, hyp, used = _ransac(sim, cu, cv, ru, rv, 1000, 0.5, 31, 2, 0, confidence=0.999, with_used=True)
        Eo, mo, co, ho, uo = PL.find_essential_philox(np.stack([cu, cv], 1), np.stack([ru, rv], 1), *K, hypotheses=1000, threshold=0.5, seed=31, frame=2,
                                                      seq=0, solver=lambda x1, x2: _solve(sim, x1, x2), confidence=0.999)
        assert (cnt, hyp, used) == (co, ho, uo) and np.array_equal(E, Eo) and np.array_equal(mask, mo)
        with np.errstate(divide="ignore"):
            need = np.log(1 - 0.999) / np.log(1 - (cnt / (e - a)) ** 5)
        assert used % 128 == 0 or used == 1000
        assert used == 1000 or (used >= need and used - 128 < max(need, 128) + 128)
        if want_rounds:
            assert used == 128 * want_rounds, (frac_bad, used, cnt, need)
        assert cnt >= 0.9 * (1 - frac_bad) * (e - a)


def test_kernel_emulation_on_tile_and_chunk_boundaries(sim):
    """Frames of 5 ... 1 500 correspondences (around the 128-candidate chunk and the 512-point tile sizes), a third of the tracks
    mismatched, several (hypotheses, confidence, grid) combinations: the kernel source == the sequential replay, bit for bit."""
    emu = load_kernel_emulation()
    rng = np.random.default_rng(123)
    sizes = [5, 6, 7, 9, 33, 127, 128, 129, 511, 512, 513, 1024, 1025, 1500, 40, 300]
    parts = []
    for i, n in enumerate(sizes):
        b = synth.make_sequence(seed=100 + i, n_frames=1, n_corr=max(n, 8), n_jitter=0.0, outlier_frac=0.2)
        parts.append([getattr(b, k)[:n].copy() for k in ("cur_u", "cur_v", "ref_u", "ref_v")])
        bad = rng.permutation(n)[: n // 3]
        parts[-1][2][bad] = rng.uniform(0, 1241, bad.size).astype(np.float32)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    arr = [np.ascontiguousarray(np.concatenate([p[j] for p in parts])) for j in range(4)]
    F = len(sizes)
    for H, conf, grid in ((260, 0.0, 5), (700, 0.999, 16), (1, 0.0, 2), (129, 0.9, 1)):
        E = np.zeros((F, 9)); mask = np.zeros(off[-1], np.uint8); cnt = np.zeros(F, np.int32); hyp = np.zeros(F, np.int32); used = np.zeros(F, np.int32)
        assert emu.fp5_emu_find_essential(F, _p(off), *(_p(a) for a in arr), *K, H, 0.5, conf, 77, None, 2, _p(E), _p(mask), _p(cnt), _p(hyp), _p(used), grid) == 0
        for f in range(F):
            a, e = off[f], off[f + 1]
            Er, mr, cr, hr, ur = _ransac(sim, *(x[a:e] for x in arr), H, 0.5, 77, f, 2, confidence=conf, with_used=True)
            assert (cr, hr, ur) == (cnt[f], hyp[f], used[f]), (H, conf, grid, sizes[f])
            assert np.array_equal(Er.reshape(-1), E[f]) and np.array_equal(mr, mask[a:e].astype(bool))


def test_kernel_numerics_agree_with_opencv_own_output(sim):
    """The reference's call -- cv2.findEssentialMat(px_cur, px_ref, K, RANSAC, 0.999, 0.5) -- stored with its recoverPose result in
    tests/golden/pose.npz (frames 0, 2, 4, 6), against the replay of the kernel's rule on the host build of its numerics with the
    same maxIters / prob / threshold: pose within the noise of the data, and (almost) the same inlier set under OpenCV's own E."""
    from oracle import extras as X
    z = np.load(os.path.join(ROOT, "tests", "golden", "pose.npz"))
    off = z["offsets"]
    ang = lambda c: np.degrees(np.arccos(np.clip(c, -1, 1)))
    for f in (0, 2, 4, 6):
        a, e = off[f], off[f + 1]
        cu, cv, ru, rv = (z[k][a:e] for k in ("cur_u", "cur_v", "ref_u", "ref_v"))
        E, mask, cnt, hyp, used = _ransac(sim, cu, cv, ru, rv, 1000, 0.5, 1, f, 0, confidence=0.999, with_used=True)
        assert used == 128 and mask.mean() > 0.9
        cur = np.stack([cu, cv], 1).astype(np.float64); ref = np.stack([ru, rv], 1).astype(np.float64)
        R, t, _, counts = X.recover_pose(E, cur, ref, *K)
        Rcv, tcv = z["R"][f].reshape(3, 3), z["t"][f]
        assert ang((np.trace(R.T @ Rcv) - 1) / 2) < 0.2 and ang(float(np.ravel(t) @ tcv)) < 1.5
        x1 = np.stack([(cur[:, 0] - K[2]) / K[0], (cur[:, 1] - K[3]) / K[1]], 1)
        x2 = np.stack([(ref[:, 0] - K[2]) / K[0], (ref[:, 1] - K[3]) / K[1]], 1)
        cv_mask = PL.sampson_inlier(z["E"][f].reshape(3, 3) / np.linalg.norm(z["E"][f]), x1, x2, (0.5 / K[0]) ** 2)
        assert (cv_mask != mask).mean() < 0.05


def test_hostile_frames_neither_hang_nor_leak_into_the_result(sim):
    """NaN / inf pixels, a frame without motion (ref == cur: the epipolar system is rank deficient), absurd magnitudes, all current
    points identical -- through the kernel source under the emulation: finite output, non-finite correspondences are never
    inliers, degenerate frames report "no model" after exhausting their hypotheses."""
    emu = load_kernel_emulation()
    z = np.load(os.path.join(ROOT, "tests", "golden", "essential.npz"))
    a, e = z["offsets"][0], z["offsets"][1]
    base = [z[k][a:e].copy() for k in ("cur_u", "cur_v", "ref_u", "ref_v")]
    f1 = [x.copy() for x in base]; f1[0][::7] = np.nan; f1[3][5::11] = np.inf
    f2 = [x.copy() for x in base]; f2[2][:] = f2[0]; f2[3][:] = f2[1]
    f3 = [np.full(50, 1e30, np.float32) for _ in range(4)]
    f4 = [x[:8].copy() for x in base]; f4[0][:] = f4[0][0]; f4[1][:] = f4[1][0]
    frames = [f1, f2, f3, f4]
    off = np.concatenate([[0], np.cumsum([len(f[0]) for f in frames])]).astype(np.int32)
    arr = [np.ascontiguousarray(np.concatenate([f[j] for f in frames]).astype(np.float32)) for j in range(4)]
    F = len(frames)
    E = np.zeros((F, 9)); mask = np.zeros(off[-1], np.uint8); cnt = np.zeros(F, np.int32); hyp = np.zeros(F, np.int32); used = np.zeros(F, np.int32)
    assert emu.fp5_emu_find_essential(F, _p(off), *(_p(x) for x in arr), *K, 256, 0.5, 0.999, 3, None, 0, _p(E), _p(mask), _p(cnt), _p(hyp), _p(used), 2) == 0
    assert np.isfinite(E).all()
    bad = ~np.isfinite(arr[0][off[0]:off[1]]) | ~np.isfinite(arr[3][off[0]:off[1]])
    m0 = mask[off[0]:off[1]].astype(bool)
    assert not (m0 & bad).any() and (m0 & ~bad).sum() >= 0.9 * (z["true_match"][a:e] & ~bad).sum() and hyp[0] >= 0 and used[0] == 128
    assert list(hyp[1:]) == [-1, -1, -1] and list(cnt[1:]) == [0, 0, 0] and list(used[1:]) == [256, 256, 256] and not mask[off[1]:].any()
