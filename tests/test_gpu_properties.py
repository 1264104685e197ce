"""Size-independent properties of the CUDA path at BASELINE's sizes (where the oracle would take minutes to hours):
determinism, shard invariance (the multi-GPU decomposition), fused == staged, Euler's formula and edge manifoldness of
every Delaunay triangulation, and the temporal filter over a fleet of sequences against the oracle's state machine."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

KITTI_LENGTHS = [4541, 1101, 4661, 801, 271, 2761, 1101, 1101, 4071, 1591, 1201]      # BASELINE configs[3]: sequences 00-10


@pytest.fixture(scope="module")
def workload():
    from mvoscalerecovery_b200 import synth
    return synth.make_sequence(seed=20261017, n_frames=1200, n_corr=2500, outlier_frac=0.10)


def _dev(engine, b):
    import torch
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(engine.device)
    return dict(offsets=t(b.offsets), cur_u=t(b.cur_u), cur_v=t(b.cur_v), ref_u=t(b.ref_u), ref_v=t(b.ref_v), poses=t(b.poses))


def _run(engine, d, maxf, lo=0, hi=None, seed=9):
    import torch
    hi = d["offsets"].numel() - 1 if hi is None else hi
    r = engine.scale_frames_from_correspondences(d["offsets"][lo:hi + 1].contiguous(), d["cur_u"], d["cur_v"], d["ref_u"], d["ref_v"],
                                                 d["poses"][lo:hi].contiguous(), max_features=maxf, frame_index0=lo, seed=seed, stats=True)
    torch.cuda.synchronize()
    return r["raw_scale"].cpu().numpy(), r["status"].cpu().numpy(), r["n_features"].cpu().numpy(), r["stats"]


def test_deterministic_and_shard_invariant(engine, workload):
    """Two runs are bit-identical; processing frame ranges separately (what fleet.frame_shards does across GPUs) gives
    bit-identical results to one launch -- the hypothesis stream is indexed by the global frame number."""
    b = workload
    d = _dev(engine, b)
    maxf = int(np.max(np.diff(b.offsets)))
    raw, st, nf, _ = _run(engine, d, maxf)
    raw2, st2, nf2, _ = _run(engine, d, maxf)
    assert np.array_equal(raw, raw2, equal_nan=True) and np.array_equal(st, st2) and np.array_equal(nf, nf2)
    from mvoscalerecovery_b200.fleet import frame_shards
    parts = [_run(engine, d, maxf, lo, hi) for lo, hi in frame_shards(b.n_frames, 3, np.diff(b.offsets))]
    assert np.array_equal(np.concatenate([p[0] for p in parts]), raw, equal_nan=True)
    assert np.array_equal(np.concatenate([p[1] for p in parts]), st)
    assert (st & 1).mean() > 0.95
    err = np.abs(raw[(st & 1) != 0] - b.true_scale[(st & 1) != 0]) / b.true_scale[(st & 1) != 0]
    assert np.median(err) < 5e-3                                    # the estimator recovers the synthetic truth


def test_fused_equals_staged_at_scale(engine, workload):
    import torch
    b = workload
    d = _dev(engine, b)
    maxf = int(np.max(np.diff(b.offsets)))
    raw, st, nf, _ = _run(engine, d, maxf)
    s1 = engine.triangulate_frames(d["offsets"], d["cur_u"], d["cur_v"], d["ref_u"], d["ref_v"], d["poses"])
    r = engine.scale_frames(d["offsets"], s1["x"], s1["y"], s1["z"], s1["u"], s1["v"], maxf, counts=s1["n_out"], seed=9)
    torch.cuda.synchronize()
    assert np.array_equal(s1["n_out"].cpu().numpy(), nf)
    assert np.array_equal(r["raw_scale"].cpu().numpy(), raw, equal_nan=True)
    assert np.array_equal(r["status"].cpu().numpy() & 0x3F, st & 0x3F)


def test_delaunay_euler_and_manifold_every_frame(engine, workload):
    """T = 2n - 2 - h (h = hull vertices, from scipy's ConvexHull) and every edge in one or two triangles, for every
    frame's ROI point set; a sample of frames additionally through the exact validator."""
    import torch
    from scipy.spatial import ConvexHull
    from oracle import exact
    b = workload
    d = _dev(engine, b)
    s1 = engine.triangulate_frames(d["offsets"], d["cur_u"], d["cur_v"], d["ref_u"], d["ref_v"], d["poses"])
    torch.cuda.synchronize()
    n_out = s1["n_out"].cpu().numpy(); U = s1["u"].cpu().numpy(); V = s1["v"].cpu().numpy()
    frames = range(0, b.n_frames, 4)
    pts = []
    for f in frames:
        a = int(b.offsets[f]); u, v = U[a:a + n_out[f]], V[a:a + n_out[f]]
        keep = v > 185
        pts.append(np.stack([u[keep], v[keep]], 1))
    off = np.zeros(len(pts) + 1, np.int32); np.cumsum([p.shape[0] for p in pts], out=off[1:])
    allp = np.concatenate(pts, 0)
    out = engine.delaunay_frames(torch.from_numpy(off).to(engine.device), torch.from_numpy(np.ascontiguousarray(allp[:, 0])).to(engine.device),
                                 torch.from_numpy(np.ascontiguousarray(allp[:, 1])).to(engine.device), int(np.max(np.diff(off))))
    torch.cuda.synchronize()
    tri = out["tri"].cpu().numpy(); ntri = out["n_tri"].cpu().numpy(); st = out["status"].cpu().numpy()
    assert not st.any()
    for i, p in enumerate(pts):
        t = tri[2 * off[i]: 2 * off[i] + ntri[i]]
        n = p.shape[0]
        h = len(ConvexHull(p.astype(np.float64)).vertices)
        assert ntri[i] == 2 * n - 2 - h, (i, ntri[i], n, h)
        e = np.sort(np.concatenate([t[:, [0, 1]], t[:, [1, 2]], t[:, [0, 2]]], 0), 1)
        _, cnt = np.unique(e, axis=0, return_counts=True)
        assert cnt.max() <= 2 and (cnt == 1).sum() == h              # boundary edges = hull edges
        assert np.all(t[:, 0] < t[:, 1]) and np.all(t[:, 1] < t[:, 2])
        assert np.all(np.lexsort((t[:, 2], t[:, 1], t[:, 0])) == np.arange(t.shape[0]))      # canonical order
        if i % 60 == 0:
            ok, msg, _ = exact.validate_delaunay(p, t, np.zeros(n, bool))
            assert ok, (i, msg)


def test_fleet_of_sequences_filter_vs_oracle(engine):
    """BASELINE configs[3] shape: 11 sequences of KITTI 00-10 lengths through ONE filter launch == the oracle's
    per-sequence state machine (slew limiter, median of 5, driver gating) and filter_10 on the same raw scales."""
    import torch
    from oracle import pipeline as P
    rng = np.random.default_rng(3)
    total = sum(KITTI_LENGTHS)
    seq_off = np.concatenate([[0], np.cumsum(KITTI_LENGTHS)]).astype(np.int32)
    truth = 0.85 + 0.3 * np.sin(np.arange(total) * 0.01) + 0.02 * rng.standard_normal(total)
    raw = truth.copy()
    jumps = rng.random(total) < 0.02
    raw[jumps] += rng.choice([-0.8, 0.9], size=jumps.sum())         # exercises the +-0.3 slew limiter
    status = np.where(rng.random(total) < 0.97, 1, 0).astype(np.uint8)      # 3 %: RANSAC did not run (state held)
    raw[status == 0] = np.nan
    move = (rng.random(total) > 0.01).astype(np.uint8)
    nfeat = np.where(rng.random(total) < 0.02, 80, 2200).astype(np.int32)
    dev = engine.device
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    out = engine.filter_sequences(t(seq_off), t(raw), t(status), t(move), t(nfeat))
    torch.cuda.synchronize()
    got, got10 = out["scale"].cpu().numpy(), out["filter10"].cpu().numpy()
    for s in range(len(KITTI_LENGTHS)):
        a, e = seq_off[s], seq_off[s + 1]
        stt = P.TemporalState(5)
        scales = [0.0]
        for f in range(a, e):
            if not move[f]:
                scales.append(0.0)
            elif nfeat[f] > P.MIN_FEATURES:
                scales.append(stt.step(raw[f], bool(status[f] & 1)))
            else:
                scales.append(scales[-1])
        ref = np.asarray(scales[1:])
        assert np.array_equal(got[a:e], ref), "sequence %d" % s
        assert np.array_equal(got10[a:e], P.filter10(ref)), "sequence %d filter_10" % s


def test_filter_under_heavy_slewing_vs_oracle(engine):
    """The slew limiter is run speculatively on the GPU (every lane starts its stretch of frames from a guessed state and the
    wrong guesses are recomputed): long slewing stretches, held states, sequences around every chunk / lane boundary -- the
    result must still be the sequential recurrence of rescale.py:168-178, bit for bit."""
    import torch
    from oracle import pipeline as P
    rng = np.random.default_rng(11)
    lens = [1, 2, 7, 8, 9, 31, 32, 33, 255, 256, 257, 1023, 1024, 1025, 3000, 0, 5000]
    total = sum(lens)
    seq_off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    raw = 0.9 + 0.05 * rng.standard_normal(total)
    level = np.cumsum(np.where(rng.random(total) < 0.03, rng.choice([-2.5, 2.5, 4.0, -4.0], size=total), 0.0))      # steps of many x 0.3
    raw = np.abs(raw + level) + 0.05
    status = np.where(rng.random(total) < 0.8, 1, 0).astype(np.uint8)
    raw[status == 0] = np.nan
    move = (rng.random(total) > 0.05).astype(np.uint8)
    nfeat = np.where(rng.random(total) < 0.1, 80, 2200).astype(np.int32)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(engine.device)
    out = engine.filter_sequences(t(seq_off), t(raw), t(status), t(move), t(nfeat))
    torch.cuda.synchronize()
    got, got10 = out["scale"].cpu().numpy(), out["filter10"].cpu().numpy()
    for s in range(len(lens)):
        a, e = seq_off[s], seq_off[s + 1]
        if a == e:
            continue
        stt = P.TemporalState(5)
        scales = [0.0]
        for f in range(a, e):
            if not move[f]:
                scales.append(0.0)
            elif nfeat[f] > P.MIN_FEATURES:
                scales.append(stt.step(raw[f], bool(status[f] & 1)))
            else:
                scales.append(scales[-1])
        ref = np.asarray(scales[1:])
        assert np.array_equal(got[a:e], ref), "sequence %d (length %d)" % (s, lens[s])
        assert np.array_equal(got10[a:e], P.filter10(ref)), "sequence %d filter_10" % s


def test_host_pipeline_with_large_frames(engine):
    """mvosr_recover_scales_host on frames beyond the shared-memory capacity: the chunks of the host pipeline run on two
    streams but share one global-memory staging buffer, so their launches must be chained; result == the device-resident
    path (fused kernel + filter)."""
    import torch
    from mvoscalerecovery_b200 import synth
    b = synth.make_sequence(seed=31, n_frames=2 * 148 + 8, n_corr=6000, outlier_frac=0.1)
    dev = engine.device
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    maxf = int(np.max(np.diff(b.offsets)))
    r = engine.scale_frames_from_correspondences(t(b.offsets), t(b.cur_u), t(b.cur_v), t(b.ref_u), t(b.ref_v), t(b.poses), max_features=maxf, seed=3)
    seq = torch.tensor([0, b.n_frames], dtype=torch.int32, device=dev)
    want = engine.filter_sequences(seq, r["raw_scale"], r["status"], t(b.move_flags), r["n_features"])["scale"].cpu().numpy()
    for _ in range(2):
        out = engine.recover_scales_host(b.offsets, b.cur_u, b.cur_v, b.ref_u, b.ref_v, b.poses, b.move_flags, max_features=maxf, seed=3)
        assert np.array_equal(out["scale"], want)
        assert np.array_equal(out["raw_scale"], r["raw_scale"].cpu().numpy(), equal_nan=True)
    assert np.isfinite(want).all() and (want > 0).mean() > 0.9


@pytest.mark.parametrize("style", ["random", "integer_pixels", "coarse_grid", "clustered", "heavy_outliers"])
def test_second_delaunay_reuse_equals_rebuild(engine, style):
    """Delaunay #2 is assembled from the stars of Delaunay #1 that lost no neighbour plus seeded rebuilds of the others
    (gstar.cuh).  It must equal, triangle for triangle, a from-scratch Delaunay of the surviving points
    (mvosr_delaunay_frames, no reuse) -- also on degenerate inputs (integer pixels: co-circular and collinear points)."""
    import torch
    from mvoscalerecovery_b200.batch import stats_to_numpy
    rng = np.random.default_rng({"random": 1, "integer_pixels": 2, "coarse_grid": 3, "clustered": 4, "heavy_outliers": 5}[style])
    F, fx, cx, cy = 24, 718.856, 607.1928, 185.2157
    f3s, f2s = [], []
    for f in range(F):
        n = int(rng.integers(300, 2600))
        u = rng.uniform(0, 1240, n); v = rng.uniform(186, 375, n)
        if style == "integer_pixels":
            u, v = np.round(u), np.round(v)
        elif style == "coarse_grid":
            u, v = 8.0 * np.round(u / 8), 186.0 + 6.0 * np.round((v - 186) / 6)
        elif style == "clustered":
            c = rng.uniform([100, 200], [1100, 360], (6, 2))
            k = rng.integers(0, 6, n)
            u = np.clip(c[k, 0] + 25 * rng.standard_normal(n), 0, 1240); v = np.clip(c[k, 1] + 12 * rng.standard_normal(n), 186, 375)
        z = 1.7 * fx / (v - cy + 1e-3) * (1 + 0.01 * rng.standard_normal(n))
        bad = rng.random(n) < (0.6 if style == "heavy_outliers" else 0.2)
        z[bad] *= rng.uniform(0.4, 0.95, bad.sum())
        f2 = np.stack([u, v], 1).astype(np.float32)
        f3 = np.stack([(u - cx) * z / fx, (v - cy) * z / fx, z], 1).astype(np.float32)
        f3s.append(f3); f2s.append(f2)
    from mvoscalerecovery_b200.batch import pack_frames
    b = pack_frames(f3s, f2s, engine.device)
    out = engine.scale_frames(b["offsets"], b["x"], b["y"], b["z"], b["u"], b["v"], b["max_features"], seed=11, debug=True)
    torch.cuda.synchronize()
    st = stats_to_numpy(out["stats"]); status = out["status"].cpu().numpy()
    dbg = {k: v.cpu().numpy() for k, v in out["debug"].items()}
    off = b["offsets"].cpu().numpy()
    pts, sel = [], []
    for f in range(F):
        if status[f] & (4 | 16 | 32) or not (status[f] & 2):
            continue                                    # no second Delaunay for this frame
        roi = f2s[f][f2s[f][:, 1] > 185]
        keep = dbg["keep"][off[f]: off[f] + roi.shape[0]].astype(bool)
        assert keep.sum() == st["n_kept"][f]
        pts.append(roi[keep]); sel.append(f)
    assert len(sel) >= F // 2
    o2 = np.zeros(len(pts) + 1, np.int32)
    np.cumsum([p.shape[0] for p in pts], out=o2[1:])
    allp = np.concatenate(pts, 0)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(engine.device)
    dt = engine.delaunay_frames(t(o2), t(allp[:, 0]), t(allp[:, 1]), int(np.max(np.diff(o2))))
    torch.cuda.synchronize()
    tri = dt["tri"].cpu().numpy(); ntri = dt["n_tri"].cpu().numpy()
    for i, f in enumerate(sel):
        got = dbg["tri2"][2 * off[f]: 2 * off[f] + st["n_tri"][f]]
        want = tri[2 * o2[i]: 2 * o2[i] + ntri[i]]
        assert st["n_tri"][f] == ntri[i], (style, f, st["n_tri"][f], ntri[i])
        assert np.array_equal(got, want), (style, f)


def test_hostile_frames_are_flagged_and_do_not_disturb_their_neighbours(engine):
    """Where the reference raises (QhullError on < 3 / collinear ROI points, src/rescale.py:124) or has no defined behaviour
    (non-finite or absurd pixel coordinates), the frame gets a status flag and a NaN raw scale; the frames around it in the same
    launch are bit-identical to a launch without the hostile frames."""
    import torch
    from mvoscalerecovery_b200 import synth, _native as N
    from mvoscalerecovery_b200.batch import pack_frames
    rng = np.random.default_rng(8)
    fx, cx, cy = 718.856, 607.1928, 185.2157

    def good(n, seed):
        r = np.random.default_rng(seed)
        u = r.uniform(0, 1240, n); v = r.uniform(186, 375, n)
        z = 1.7 * fx / (v - cy) * (1 + 0.005 * r.standard_normal(n))
        return np.stack([(u - cx) * z / fx, (v - cy) * z / fx, z], 1).astype(np.float32), np.stack([u, v], 1).astype(np.float32)

    g0, g1, g2 = good(900, 1), good(1500, 2), good(400, 3)
    nan2 = g0[1].copy(); nan2[::7, 0] = np.nan
    far2 = g0[1].copy(); far2[5, 0] = 5000.0
    inf3 = g0[0].copy(); inf3[::5] = np.inf                                 # non-finite 3-D points, valid pixels
    line2 = np.stack([np.linspace(10, 1200, 300), np.full(300, 250.0)], 1).astype(np.float32)
    dup2 = np.tile(np.array([[300.0, 300.0]], np.float32), (200, 1))
    sky2 = g0[1].copy(); sky2[:, 1] = 100.0                                  # everything above the ROI row
    f3 = [g0[0], g0[0], g0[0], g1[0], inf3, g0[0][:300], g0[0][:200], g0[0], g0[0][:2], np.zeros((0, 3), np.float32), g2[0]]
    f2 = [g0[1], nan2, far2, g1[1], g0[1], line2, dup2, sky2, g0[1][:2], np.zeros((0, 2), np.float32), g2[1]]
    b = pack_frames(f3, f2, engine.device)
    out = engine.scale_frames(b["offsets"], b["x"], b["y"], b["z"], b["u"], b["v"], b["max_features"], seed=4)
    torch.cuda.synchronize()
    st = out["status"].cpu().numpy(); raw = out["raw_scale"].cpu().numpy()
    assert st[0] & N.ST_UPDATED and st[3] & N.ST_UPDATED and st[10] & N.ST_UPDATED
    assert st[1] & N.ST_BAD_INPUT and st[2] & N.ST_BAD_INPUT and np.isnan(raw[1]) and np.isnan(raw[2])
    assert st[4] & N.ST_BAD_INPUT and not (st[4] & N.ST_UPDATED) and np.isnan(raw[4])       # non-finite 3-D coordinates in the ROI: flagged, state held
    for f in (5, 6, 7, 8, 9):
        assert st[f] & N.ST_FEW_ROI and np.isnan(raw[f]), (f, st[f])
    # the good frames alone, same frame indices for the hypothesis stream
    for f, (a3, a2) in ((0, g0), (3, g1), (10, g2)):
        bb = pack_frames([a3], [a2], engine.device)
        o = engine.scale_frames(bb["offsets"], bb["x"], bb["y"], bb["z"], bb["u"], bb["v"], bb["max_features"], frame_index0=f, seed=4)
        assert float(o["raw_scale"].cpu().numpy()[0]) == raw[f] and int(o["status"].cpu().numpy()[0]) == st[f]
    # features on a plane THROUGH the camera centre (Y = Z / 4, dyadic coordinates: every product below is exact): every triangle's
    # vertex matrix is singular, the reference raises LinAlgError in np.matrix(P).I (rescale.py:79) -> MVOSR_ST_SINGULAR, no
    # RANSAC, NaN raw scale, state held.  Depth falls strictly with the pixel row, so the graph check keeps everything.
    k = np.arange(40)
    Zs = 50.0 - 0.5 * k
    s3 = np.stack([np.where(k % 2 == 0, 1.0, -2.0) * (1 + k % 5), Zs / 4, Zs], 1).astype(np.float32)
    s2 = np.stack([(37.0 * k * k + 11.0 * k) % 1200.0 + 10.0, 200.0 + 3.0 * k], 1).astype(np.float32)
    bb = pack_frames([s3], [s2], engine.device)
    o = engine.scale_frames(bb["offsets"], bb["x"], bb["y"], bb["z"], bb["u"], bb["v"], bb["max_features"], seed=4)
    so = int(o["status"].cpu().numpy()[0])
    assert so & N.ST_SINGULAR and not (so & N.ST_UPDATED) and np.isnan(float(o["raw_scale"].cpu().numpy()[0])), so
    # the filter holds its state over flagged frames (reference: uncaught exception -> undefined; here: hold)
    seq = torch.tensor([0, len(f3)], dtype=torch.int32, device=engine.device)
    nf = torch.tensor([a.shape[0] for a in f3], dtype=torch.int32, device=engine.device)
    sc = engine.filter_sequences(seq, out["raw_scale"], out["status"], None, nf)["scale"].cpu().numpy()
    assert np.isfinite(sc).all()


def test_fleet_host_call_equals_device_path(engine):
    """mvosr_recover_fleet_host: several sequences in one host-buffer call (frame counters, Philox sequence ids and the temporal
    filter restart at every sequence) == per-sequence fused launches + one filter launch over the sequence offsets."""
    import torch
    from mvoscalerecovery_b200 import synth
    lens = [310, 7, 0, 150]
    parts = [synth.make_sequence(seed=44, n_frames=L, n_corr=700, seq=5 + s, outlier_frac=0.15, still_every=13) for s, L in enumerate(lens)]
    so = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    off = np.concatenate([[0]] + [p.offsets[1:].astype(np.int64) + sum(int(q.offsets[-1]) for q in parts[:i]) for i, p in enumerate(parts)]).astype(np.int32)
    cat = lambda k: np.concatenate([getattr(p, k) for p in parts])
    cu, cv, ru, rv, poses, move = (cat(k) for k in ("cur_u", "cur_v", "ref_u", "ref_v", "poses", "move_flags"))
    maxf = int(np.max(np.diff(off)))
    dev = engine.device
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    raws, sts, nfs = [], [], []
    for s, p in enumerate(parts):
        if p.n_frames == 0:
            continue
        r = engine.scale_frames_from_correspondences(t(p.offsets), t(p.cur_u), t(p.cur_v), t(p.ref_u), t(p.ref_v), t(p.poses),
                                                     max_features=maxf, frame_index0=0, seq_id=5 + s, seed=21)
        raws.append(r["raw_scale"]); sts.append(r["status"]); nfs.append(r["n_features"])
    want = engine.filter_sequences(t(so), torch.cat(raws), torch.cat(sts), t(move), torch.cat(nfs))["scale"].cpu().numpy()
    out = engine.recover_scales_host(off, cu, cv, ru, rv, poses, move, max_features=maxf, seq_id=5, seed=21, seq_offsets=so)
    assert np.array_equal(out["scale"], want)
    assert np.array_equal(out["raw_scale"], torch.cat(raws).cpu().numpy(), equal_nan=True)
    assert np.array_equal(out["status"], torch.cat(sts).cpu().numpy())


@pytest.mark.parametrize("iters,stop", [(100, 0), (37, 1), (300, 0), (1, 1)])
def test_ransac_configurations_vs_oracle(iters, stop):
    """BASELINE configs[4] (the RANSAC sweep): other hypothesis counts and "evaluate all hypotheses" (no early stop) through the
    fused kernel == the oracle's sequential run_ransac over the same Philox stream -- chosen hypothesis, inlier count,
    hypotheses used and raw scale, frame by frame."""
    import torch
    from mvoscalerecovery_b200 import synth
    from mvoscalerecovery_b200.batch import ScaleRecovery, stats_to_numpy
    from oracle import pipeline as P
    eng = ScaleRecovery(absolute_reference=1.7, ransac_iterations=iters, ransac_stop_at_goal=stop)
    b = synth.make_sequence(seed=61, n_frames=6, n_corr=900, outlier_frac=0.3)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(eng.device)
    s1 = eng.triangulate_frames(t(b.offsets), t(b.cur_u), t(b.cur_v), t(b.ref_u), t(b.ref_v), t(b.poses))
    maxf = int(np.max(np.diff(b.offsets)))
    out = eng.scale_frames(t(b.offsets), s1["x"], s1["y"], s1["z"], s1["u"], s1["v"], maxf, counts=s1["n_out"], seed=17, seq_id=2)
    torch.cuda.synchronize()
    st = stats_to_numpy(out["stats"]); raw = out["raw_scale"].cpu().numpy(); n_out = s1["n_out"].cpu().numpy()
    for f in range(b.n_frames):
        a = int(b.offsets[f]); m = int(n_out[f])
        f3 = np.stack([s1[k][a:a + m].cpu().numpy() for k in "xyz"], 1).astype(np.float64)
        f2 = np.stack([s1[k][a:a + m].cpu().numpy() for k in "uv"], 1).astype(np.float64)
        rec = P.feature_selection(f3, f2)
        rr = P.ransac_plane(rec["point_selected"], 17, f, seq=2, max_iterations=iters, stop_at_goal=bool(stop))
        assert (st["best_hyp"][f], st["best_ic"][f], st["hyps_used"][f]) == (rr["best_hyp"], rr["ic"], rr["hyps_used"]), (f, iters, stop)
        np.testing.assert_allclose(raw[f], 1.7 / P.height_from_model(rr["model"]), rtol=1e-9)
    eng.close()


@pytest.mark.parametrize("density", ["ground", "clustered"])
def test_other_feature_densities_vs_oracle(engine, density):
    """The bench's perspective (SURVEY 8d ground generator) and clustered feature distributions at the headline size (2 500
    correspondences, ~2 000 ROI features), full pipeline against the oracle frame by frame: survivor count, triangle count of
    Delaunay #2, vertex-list length, chosen hypothesis, inlier count, hypotheses used exactly, raw scale to 1e-9 -- the strip index
    and the star paths see very uneven candidate blocks here (the round-1 uniform grid sent most of these stars to its slow paths)."""
    import torch
    from mvoscalerecovery_b200 import synth
    from mvoscalerecovery_b200.batch import stats_to_numpy
    from oracle import pipeline as P
    b = synth.make_sequence(seed=4242, n_frames=4, n_corr=2500, outlier_frac=0.10, density=density)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(engine.device)
    s1 = engine.triangulate_frames(t(b.offsets), t(b.cur_u), t(b.cur_v), t(b.ref_u), t(b.ref_v), t(b.poses))
    maxf = int(np.max(np.diff(b.offsets)))
    out = engine.scale_frames(t(b.offsets), s1["x"], s1["y"], s1["z"], s1["u"], s1["v"], maxf, counts=s1["n_out"], seed=5, seq_id=1, stats=True)
    fused = engine.scale_frames_from_correspondences(t(b.offsets), t(b.cur_u), t(b.cur_v), t(b.ref_u), t(b.ref_v), t(b.poses),
                                                     max_features=maxf, seed=5, seq_id=1)
    torch.cuda.synchronize()
    st = stats_to_numpy(out["stats"]); raw = out["raw_scale"].cpu().numpy(); n_out = s1["n_out"].cpu().numpy()
    assert np.array_equal(raw, fused["raw_scale"].cpu().numpy(), equal_nan=True)
    assert (out["status"].cpu().numpy() & 1).all()
    for f in range(b.n_frames):
        a = int(b.offsets[f]); m = int(n_out[f])
        f3 = np.stack([s1[k][a:a + m].cpu().numpy() for k in "xyz"], 1).astype(np.float64)
        f2 = np.stack([s1[k][a:a + m].cpu().numpy() for k in "uv"], 1).astype(np.float64)
        rec = P.frame_raw_scale(f3, f2, 5, f, 1, absolute_reference=1.7)
        assert st["n_kept"][f] == int(rec["keep"].sum()), (density, f)
        assert st["n_tri"][f] == rec["tri2"].shape[0], (density, f)
        assert 3 * st["n_valid"][f] == rec["n_sel"], (density, f)
        assert (st["best_ic"][f], st["hyps_used"][f]) == (rec["ic"], rec["hyps_used"]), (density, f)
        np.testing.assert_allclose(raw[f], rec["raw_scale"], rtol=1e-9)


def _hub_points(rng, spokes, n_bg, integer=False):
    """A hub: one feature in the middle of a ring of `spokes` features (its Delaunay star has that many neighbours) over a random
    background that stays outside the ring.  integer=True: the ring on an exact circle of integer radius around an integer centre is not
    possible in general, so the ring points are rounded -- ties and near-ties among the spokes go through the exact predicates."""
    c = np.array([620.0, 285.0])
    ang = (np.arange(spokes) + rng.uniform(-0.2, 0.2, spokes)) * (2 * np.pi / spokes)          # (a spoke set back by more than the sag of its
    r = 70.0 + rng.uniform(-0.002, 0.002, spokes)                                                  # neighbours' chord would lose its edge to the centre)
    ring = c + np.stack([r * np.cos(ang), r * np.sin(ang)], 1)
    bg = np.stack([rng.uniform(0, 1240, 4 * n_bg), rng.uniform(186, 375, 4 * n_bg)], 1)
    bg = bg[np.linalg.norm(bg - c, axis=1) > 78.0][:n_bg]
    pts = np.concatenate([c[None], ring, bg], 0)
    if integer:
        pts = np.round(pts)
        pts = pts[np.sort(np.unique(pts, axis=0, return_index=True)[1])]
    return pts[rng.permutation(pts.shape[0])].astype(np.float32)


@pytest.mark.parametrize("spokes", [20, 40, 100, 230])
def test_hub_stars_any_degree_vs_qhull(engine, spokes):
    """A star of more than 32 neighbours used to end its frame with MVOSR_ST_OVERFLOW (fb_build keeps a star on the 32 lanes of a warp);
    Qhull has no such limit.  Level 5 (hub_star) builds such stars in memory: triangulations with hubs of 20 ... 230 spokes equal
    Qhull's, and the rounded (tie-laden) variants equal the exact-arithmetic oracle's."""
    import torch
    from scipy.spatial import Delaunay
    from oracle import ref_harness as H
    rng = np.random.default_rng(1000 + spokes)
    sets = [_hub_points(rng, spokes, 400), _hub_points(rng, spokes, 1500), _hub_points(rng, spokes, 60)]
    off = np.zeros(len(sets) + 1, np.int32)
    np.cumsum([p.shape[0] for p in sets], out=off[1:])
    allp = np.concatenate(sets, 0)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(engine.device)
    out = engine.delaunay_frames(t(off), t(allp[:, 0]), t(allp[:, 1]), int(np.max(np.diff(off))))
    torch.cuda.synchronize()
    tri = out["tri"].cpu().numpy(); ntri = out["n_tri"].cpu().numpy(); st = out["status"].cpu().numpy()
    for i, p in enumerate(sets):
        assert st[i] == 0, (spokes, i, st[i])
        want = H.canonicalise(Delaunay(p.astype(np.float64)).simplices)
        got = tri[2 * off[i]: 2 * off[i] + ntri[i]]
        assert ntri[i] == want.shape[0] and np.array_equal(got, want), (spokes, i)
        deg = np.bincount(want.reshape(-1), minlength=p.shape[0]).max()
        assert deg >= 0.9 * spokes, (spokes, deg)


def test_hub_frames_full_pipeline_vs_oracle(engine):
    """The same through the whole estimator (vote pass, ring store, second Delaunay, planes, RANSAC): frames whose road features
    contain a 60-spoke hub against the oracle, and the status byte stays clean."""
    import torch
    from mvoscalerecovery_b200.batch import pack_frames, stats_to_numpy
    from oracle import pipeline as P
    rng = np.random.default_rng(77)
    fx, cx, cy = 718.856, 607.1928, 185.2157
    f3s, f2s = [], []
    for f in range(3):
        p = _hub_points(rng, 60, 900).astype(np.float64)
        u, v = p[:, 0], p[:, 1]
        z = 1.7 * fx / (v - cy + 1e-3) * (1 + 0.01 * rng.standard_normal(u.shape[0]))
        bad = rng.random(u.shape[0]) < 0.15
        z[bad] *= rng.uniform(0.4, 0.95, bad.sum())
        f2s.append(np.stack([u, v], 1).astype(np.float32)); f3s.append(np.stack([(u - cx) * z / fx, (v - cy) * z / fx, z], 1).astype(np.float32))
    b = pack_frames(f3s, f2s, engine.device)
    out = engine.scale_frames(b["offsets"], b["x"], b["y"], b["z"], b["u"], b["v"], b["max_features"], seed=3, seq_id=0, stats=True)
    torch.cuda.synchronize()
    st = stats_to_numpy(out["stats"]); raw = out["raw_scale"].cpu().numpy(); status = out["status"].cpu().numpy()
    for f in range(3):
        assert not (status[f] & 32), (f, status[f])
        rec = P.frame_raw_scale(f3s[f].astype(np.float64), f2s[f].astype(np.float64), 3, f, 0, absolute_reference=1.7)
        assert st["n_kept"][f] == int(rec["keep"].sum()) and st["n_tri"][f] == rec["tri2"].shape[0], f
        assert (st["best_ic"][f], st["hyps_used"][f], 3 * st["n_valid"][f]) == (rec["ic"], rec["hyps_used"], rec["n_sel"]), f
        np.testing.assert_allclose(raw[f], rec["raw_scale"], rtol=1e-9)
