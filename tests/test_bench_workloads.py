"""bench.py's host-side workload construction (what every rank processes) -- no GPU: the fleet workload (BASELINE configs[3])
must cover every frame of every sequence exactly once across the ranks, with each rank's batch equal to the corresponding
slices of the per-sequence data; the weak-scaling workloads put sequence `rank` on rank `rank`."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                                     # noqa: E402
from mvoscalerecovery_b200 import synth                          # noqa: E402


def test_fleet_pieces_cover_every_frame_once():
    lengths = (13, 4, 1, 22, 7)
    full = [synth.make_sequence(seed=bench.SEED, n_frames=L, n_corr=60, seq=s, outlier_frac=0.10) for s, L in enumerate(lengths)]
    for world in (1, 2, 3, 5):
        seen = {s: np.zeros(L, int) for s, L in enumerate(lengths)}
        total = 0
        for rank in range(world):
            wl = bench.build_rank_workload("fleet", rank, world, features=60, lengths=lengths)
            assert wl["total_frames"] == sum(lengths) and wl["shards"][rank][1] - wl["shards"][rank][0] == wl["n_frames"]
            assert list(wl["seq_off_host"]) == list(np.concatenate([[0], np.cumsum(lengths)]))
            b, f0 = wl["batch"], 0
            assert b.n_frames == wl["n_frames"] == sum(e - a for _, a, e in wl["pieces"])
            for sq, a, e in wl["pieces"]:
                seen[sq][a:e] += 1
                ref = full[sq]
                lo, hi = b.offsets[f0], b.offsets[f0 + (e - a)]
                assert np.array_equal(b.cur_u[lo:hi], ref.cur_u[ref.offsets[a]:ref.offsets[e]])
                assert np.array_equal(b.poses[f0:f0 + (e - a)], ref.poses[a:e])
                assert np.array_equal(np.diff(b.offsets[f0:f0 + (e - a) + 1]), np.diff(ref.offsets[a:e + 1]))
                f0 += e - a
            total += b.n_frames
        assert total == sum(lengths) and all((v == 1).all() for v in seen.values())


def test_weak_scaling_workload_is_one_sequence_per_rank():
    for rank in (0, 2):
        wl = bench.build_rank_workload("kitti00", rank, 4, frames=5, features=80)
        assert wl["pieces"] == [(rank, 0, 5)] and wl["total_frames"] == 20 and list(wl["seq_off_host"]) == [0, 5, 10, 15, 20]
        ref = synth.make_sequence(seed=bench.SEED, n_frames=5, n_corr=80, seq=rank, outlier_frac=0.10)
        assert np.array_equal(wl["batch"].cur_u, ref.cur_u) and wl["shards"][rank] == (5 * rank, 5 * rank + 5)
    assert {"kitti00", "dense", "fleet", "kitti00-ground", "kitti00-clustered"} <= set(bench.WORKLOADS) and sum(bench.KITTI_LENGTHS) == 23201
