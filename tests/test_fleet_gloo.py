"""Multi-GPU plumbing on CPU: frame-range sharding and the one all-gather of raw per-frame results, world_size 2,
gloo backend (the GPU box runs the same code over NCCL)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_frame_shards_cover_and_balance():
    from mvoscalerecovery_b200.fleet import frame_shards
    lens = [4541, 1101, 4661, 801, 271, 2761, 1101, 1101, 4071, 1591, 1201]        # KITTI 00-10 (BASELINE configs[3])
    total = sum(lens)
    for w in (1, 2, 4, 8):
        sh = frame_shards(total, w)
        assert sh[0][0] == 0 and sh[-1][1] == total and all(a[1] == b[0] for a, b in zip(sh, sh[1:]))
        assert max(e - s for s, e in sh) - min(e - s for s, e in sh) <= 1
    wts = np.concatenate([np.full(n, 1000 + 100 * i) for i, n in enumerate(lens)])
    sh = frame_shards(total, 4, wts)
    loads = [wts[s:e].sum() for s, e in sh]
    assert sh[0][0] == 0 and sh[-1][1] == total and max(loads) / min(loads) < 1.01
    with pytest.raises(ValueError):
        frame_shards(10, 0)


def _rank_main(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mvoscalerecovery_b200.fleet import frame_shards, gather_results
    ok = True
    for total in (23, 24):                      # uneven shards (11 + 12: padded) and even ones (12 + 12: the unpadded fast path)
        shards = frame_shards(total, world)
        s, e = shards[rank]
        g = torch.Generator().manual_seed(5)
        raw_all = torch.rand(total, dtype=torch.float64, generator=g) + 0.5
        raw_all[3] = float("nan")
        st_all = torch.randint(0, 128, (total,), dtype=torch.uint8, generator=g)
        nf_all = torch.randint(0, 3000, (total,), dtype=torch.int32, generator=g)
        raw, st, nf = gather_results(raw_all[s:e].clone(), st_all[s:e].clone(), nf_all[s:e].clone(), shards)
        ok = ok and torch.equal(torch.nan_to_num(raw, nan=-1.0), torch.nan_to_num(raw_all, nan=-1.0)) and torch.equal(st, st_all) and torch.equal(nf, nf_all)
        ok = ok and raw.is_contiguous() and st.dtype == torch.uint8 and nf.dtype == torch.int32
    with open(os.path.join(tmp, "rank%d" % rank), "w") as f:
        f.write("ok" if ok else "mismatch")
    dist.destroy_process_group()


def test_gather_results_world_size_2_gloo(tmp_path):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_rank_main, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(os.path.join(str(tmp_path), "rank%d" % r)).read() == "ok"


def _essential_rank_main(rank, world, port, tmp):
    """Five-point RANSAC sharded by frame range: every rank estimates its frames addressed by their GLOBAL index (frame_index of
    mvosr_find_essential_frames; here the host build of the kernel's numerics stands in for the device), one gather, and the
    result equals the unsharded run -- the sample stream is a function of (seed, sequence, frame), not of the shard."""
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mvoscalerecovery_b200.fleet import frame_shards, gather_results
    import test_five_point_host_sim as T
    L = T.load_host_sim()
    z = np.load(os.path.join(ROOT, "tests", "golden", "essential.npz"))
    off = z["offsets"]
    F = len(off) - 1
    run = lambda f: T._ransac(L, *(z[k][off[f]:off[f + 1]] for k in ("cur_u", "cur_v", "ref_u", "ref_v")), 150, 0.5, 9, f, 2, confidence=0.999, with_used=True)
    shards = frame_shards(F, world)
    s, e = shards[rank]
    mine = [run(f) for f in range(s, e)]
    cnt, used, hyp = gather_results(torch.tensor([float(m[2]) for m in mine], dtype=torch.float64), torch.tensor([m[4] for m in mine], dtype=torch.uint8),
                                    torch.tensor([m[3] for m in mine], dtype=torch.int32), shards)
    whole = [run(f) for f in range(F)]
    ok = [int(c) for c in cnt] == [w[2] for w in whole] and hyp.tolist() == [w[3] for w in whole] and used.tolist() == [w[4] for w in whole]
    with open(os.path.join(tmp, "ess_rank%d" % rank), "w") as f:
        f.write("ok" if ok else "mismatch")
    dist.destroy_process_group()


def test_sharded_essential_estimation_world_size_2_gloo(tmp_path):
    world = 2
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_essential_rank_main, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(os.path.join(str(tmp_path), "ess_rank%d" % r)).read() == "ok"


def test_slot_map_and_frame_tables():
    from mvoscalerecovery_b200.fleet import frame_shards, frame_tables, slot_map
    lens = [7, 3, 1, 9]
    starts = np.concatenate([[0], np.cumsum(lens)])
    shards = frame_shards(int(starts[-1]), 3)
    slot = slot_map(shards)
    max_len = max(e - s for s, e in shards)
    assert len(set(slot.tolist())) == int(starts[-1])                              # injective
    for r, (s, e) in enumerate(shards):
        assert slot[s:e].tolist() == list(range(r * max_len, r * max_len + e - s))
    sq, fi = frame_tables(starts, 0, int(starts[-1]))
    assert sq.tolist() == sum([[s] * n for s, n in enumerate(lens)], []) and fi.tolist() == sum([list(range(n)) for n in lens], [])
    lo, hi = shards[1]
    sq1, fi1 = frame_tables(starts, lo, hi)
    assert np.array_equal(sq1, sq[lo:hi]) and np.array_equal(fi1, fi[lo:hi])       # a frame's stream does not depend on the cut


def _record_rank_main(rank, world, port, tmp):
    """RecordExchange over gloo: every rank writes its records into its block, ONE in-place all-gather, and frame f of the global
    order is found at slot[f] -- for uneven shards (padded blocks) and even ones."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mvoscalerecovery_b200.fleet import RecordExchange, frame_shards, records_to_numpy
    ok = True
    dt = np.dtype([("raw_scale", np.float64), ("n_features", np.int32), ("status", np.uint8), ("pad", np.uint8, (3,))])
    for total in (23, 24):
        shards = frame_shards(total, world)
        s, e = shards[rank]
        rng = np.random.default_rng(7)
        allrec = np.zeros(total, dt)
        allrec["raw_scale"] = rng.uniform(0.5, 1.5, total); allrec["raw_scale"][3] = np.nan
        allrec["n_features"] = rng.integers(0, 3000, total); allrec["status"] = rng.integers(0, 128, total)
        ex = RecordExchange(shards, rank, torch.device("cpu"))
        ex.mine[: e - s] = torch.from_numpy(allrec[s:e].view(np.uint8).reshape(-1, 16).copy())
        got = records_to_numpy(ex.gather())[ex.slot.numpy()]
        ok = ok and got.tobytes() == allrec.tobytes()
    with open(os.path.join(tmp, "rec_rank%d" % rank), "w") as f:
        f.write("ok" if ok else "mismatch")
    dist.destroy_process_group()


def test_record_exchange_world_size_2_gloo(tmp_path):
    world = 2
    port = 33500 + (os.getpid() % 2000)
    mp.spawn(_record_rank_main, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(os.path.join(str(tmp_path), "rec_rank%d" % r)).read() == "ok"
