"""The fleet path on the GPU (BASELINE configs[3]; reference loop: src/main_offline.py:57-88 run once per sequence): one shard
launch over a frame range that spans sequences, 16-byte records, the filter reading them through the slot map -- against the
per-sequence entry points, bit for bit.  The N>1 exchange itself is covered on CPU (tests/test_fleet_gloo.py, gloo)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _fleet(lens, n_corr=700, seq0=5, seed_data=44):
    from mvoscalerecovery_b200 import synth
    parts = [synth.make_sequence(seed=seed_data, n_frames=L, n_corr=n_corr, seq=seq0 + s, outlier_frac=0.15, still_every=13, n_jitter=0.3)
             for s, L in enumerate(lens)]
    so = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    off = np.concatenate([[0]] + [p.offsets[1:].astype(np.int64) + sum(int(q.offsets[-1]) for q in parts[:i]) for i, p in enumerate(parts)]).astype(np.int32)
    cat = lambda k: np.concatenate([getattr(p, k) for p in parts])
    return parts, so, off, {k: cat(k) for k in ("cur_u", "cur_v", "ref_u", "ref_v", "poses", "move_flags")}


def _per_sequence(engine, parts, so, move, maxf, seq0, seed):
    import torch
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(engine.device)
    raws, sts, nfs = [], [], []
    for s, p in enumerate(parts):
        if p.n_frames == 0:
            continue
        r = engine.scale_frames_from_correspondences(t(p.offsets), t(p.cur_u), t(p.cur_v), t(p.ref_u), t(p.ref_v), t(p.poses),
                                                     max_features=maxf, frame_index0=0, seq_id=seq0 + s, seed=seed)
        raws.append(r["raw_scale"]); sts.append(r["status"]); nfs.append(r["n_features"])
    raw, st, nf = torch.cat(raws), torch.cat(sts), torch.cat(nfs)
    flt = engine.filter_sequences(t(so), raw, st, t(move), nf)
    return raw.cpu().numpy(), st.cpu().numpy(), nf.cpu().numpy(), flt["scale"].cpu().numpy(), flt["filter10"].cpu().numpy()


def test_shard_launch_equals_per_sequence_launches(engine):
    """mvosr_scale_shard_from_correspondences over ranges that cut through sequences, in any processing order, writes the records
    the per-sequence launches produce; mvosr_filter_records over those records (scattered by a slot map, as after a padded
    all-gather) == mvosr_filter_sequences."""
    import torch
    from mvoscalerecovery_b200 import fleet
    lens = [310, 7, 0, 150]
    parts, so, off, a = _fleet(lens)
    maxf = int(np.max(np.diff(off)))
    raw, st, nf, want, want10 = _per_sequence(engine, parts, so, a["move_flags"], maxf, 5, 21)
    dev = engine.device
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    F = int(so[-1])
    shards = [(0, 100), (100, 330), (330, F)]                   # the second one spans three sequences
    max_len = max(e - s for s, e in shards)
    buf = torch.zeros(len(shards), max_len, 16, dtype=torch.uint8, device=dev)
    d = {k: t(a[k]) for k in ("cur_u", "cur_v", "ref_u", "ref_v", "poses")}
    d_off = t(off)
    rng = np.random.default_rng(3)
    for r, (lo, hi) in enumerate(shards):
        fseq, fidx = fleet.frame_tables(so, lo, hi)
        order = rng.permutation(hi - lo).astype(np.int32)
        # a shard's batch starts at its own offset 0 on a real rank; here it is a view of the fleet's arrays (offsets are absolute)
        engine.scale_shard_from_correspondences(d_off[lo:hi + 1], d["cur_u"], d["cur_v"], d["ref_u"], d["ref_v"], d["poses"][lo:hi], maxf,
                                                buf[r], frame_seq=t(fseq + 5), frame_index=t(fidx), order=t(order), seed=21)
    slot = t(fleet.slot_map(shards))
    out = engine.filter_records(t(so), buf, slot, t(a["move_flags"]))
    torch.cuda.synchronize()
    rec = fleet.records_to_numpy(buf)[fleet.slot_map(shards)]
    assert np.array_equal(rec["raw_scale"], raw, equal_nan=True)
    assert np.array_equal(rec["status"], st) and np.array_equal(rec["n_features"], nf)
    assert np.array_equal(out["scale"].cpu().numpy(), want) and np.array_equal(out["filter10"].cpu().numpy(), want10)


def test_shard_runner_device_and_host_steps(engine):
    """fleet.ShardRunner (world size 1): step() and step_host() (chunked H2D on a copy stream, two compute streams) give the
    per-sequence result; the seq/frame tables make the frame-index restart at every sequence."""
    import torch
    from mvoscalerecovery_b200 import fleet
    lens = [600, 90, 410]
    parts, so, off, a = _fleet(lens, n_corr=500, seq0=0, seed_data=45)
    maxf = int(np.max(np.diff(off)))
    raw, st, nf, want, want10 = _per_sequence(engine, parts, so, a["move_flags"], maxf, 0, 9)
    host = dict(offsets=off, **{k: a[k] for k in ("cur_u", "cur_v", "ref_u", "ref_v", "poses")})
    run = fleet.ShardRunner(engine, host, [(0, int(so[-1]))], 0, so, seed=9, move_flags=a["move_flags"], host_chunks=4)
    out = run.step()
    torch.cuda.synchronize()
    assert np.array_equal(out["scale"].cpu().numpy(), want) and np.array_equal(out["filter10"].cpu().numpy(), want10)
    rec = fleet.records_to_numpy(run.ex.buffer)
    assert np.array_equal(rec["raw_scale"], raw, equal_nan=True) and np.array_equal(rec["status"], st)
    run.ex.buffer.zero_()
    assert len(run.chunks) == 4
    got = run.step_host()
    assert np.array_equal(got.numpy(), want)


def test_host_entry_rejects_malformed_buffers(engine):
    """the host-buffer entry points take raw pointers -- wrong dtype, strided views, short arrays and non-monotone
    offsets must be refused before the C side reinterprets them."""
    from mvoscalerecovery_b200 import synth
    b = synth.make_sequence(seed=2, n_frames=4, n_corr=300)
    args = [b.offsets, b.cur_u, b.cur_v, b.ref_u, b.ref_v, b.poses, b.move_flags]
    ok = engine.recover_scales_host(*args, seed=1)
    assert ok["scale"].shape == (4,)
    bad = list(args); bad[0] = b.offsets.astype(np.int64)
    with pytest.raises(ValueError):
        engine.recover_scales_host(*bad, seed=1)
    bad = list(args); bad[1] = b.cur_u.astype(np.float64)
    with pytest.raises(ValueError):
        engine.recover_scales_host(*bad, seed=1)
    bad = list(args); bad[2] = np.concatenate([b.cur_v, b.cur_v])[::2]
    with pytest.raises(ValueError):
        engine.recover_scales_host(*bad, seed=1)
    bad = list(args); bad[3] = b.ref_u[:-5]
    with pytest.raises(ValueError):
        engine.recover_scales_host(*bad, seed=1)
    bad = list(args); o = b.offsets.copy(); o[2] = o[1] - 1; bad[0] = o
    with pytest.raises(ValueError):
        engine.recover_scales_host(*bad, seed=1)
    import torch
    with pytest.raises(ValueError):                              # optional device tensors are checked too
        t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(engine.device)
        engine.scale_frames_from_correspondences(t(b.offsets), t(b.cur_u), t(b.cur_v), t(b.ref_u), t(b.ref_v), t(b.poses), 400,
                                                 e_mask=torch.ones(int(b.offsets[-1]), dtype=torch.bool, device=engine.device))
