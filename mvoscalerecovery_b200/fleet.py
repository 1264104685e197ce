"""Multi-GPU plumbing: shard a fleet of sequences by frame range, gather raw scales, filter.

Frames are independent up to the temporal filter, so the data path has no
collective: every rank runs stages 1-5 on its contiguous frame range.  Only the per-frame raw
results (raw scale f64, status, feature count: 24 B/frame packed as 3 float64) are exchanged with
ONE all-gather; then the strictly sequential stage 6 (slew limiter + deque median,
src/rescale.py:168-178, and the gating of src/main_offline.py:57-88) runs on the full vector.
A 10-frame halo would be exact for the windowed medians but not for the slew limiter, whose
memory is unbounded; gather-then-filter is exact and the payload is a few hundred KB.

Works with any torch.distributed backend: NCCL on the GPU box, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def frame_shards(n_frames_total: int, world_size: int, weights: Sequence[float] = None) -> List[Tuple[int, int]]:
    """Contiguous [start, end) frame ranges, one per rank.  With ``weights`` (e.g. features per
    frame) the ranges balance the summed weight instead of the frame count."""
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    if weights is None:
        cuts = [(n_frames_total * r) // world_size for r in range(world_size + 1)]
    else:
        w = np.asarray(weights, dtype=np.float64)
        if w.shape[0] != n_frames_total:
            raise ValueError("weights must have one entry per frame")
        c = np.concatenate([[0.0], np.cumsum(w)])
        targets = c[-1] * np.arange(world_size + 1) / world_size
        cuts = [int(np.searchsorted(c, t, side="left")) for t in targets]
        cuts[0], cuts[-1] = 0, n_frames_total
        for i in range(1, len(cuts)):
            cuts[i] = max(cuts[i], cuts[i - 1])
    return [(cuts[r], cuts[r + 1]) for r in range(world_size)]


def pack_results(raw_scale: torch.Tensor, status: torch.Tensor, n_features: torch.Tensor, pad_to: int) -> torch.Tensor:
    """(L,) f64 / u8 / i32 -> (pad_to, 3) f64 (status and counts are exactly representable)."""
    L = raw_scale.numel()
    packed = torch.stack([raw_scale, status.to(torch.float64), n_features.to(torch.float64)], 1)
    if L == pad_to:
        return packed
    out = torch.zeros(pad_to, 3, dtype=torch.float64, device=raw_scale.device)
    out[:L] = packed
    return out


def gather_results(raw_scale, status, n_features, shards: List[Tuple[int, int]], group=None):
    """All-gather the per-frame raw results of every rank's shard; returns full-length
    (raw_scale f64, status u8, n_features i32) in global frame order on every rank."""
    world = len(shards)
    if world == 1:                                            # nothing to exchange: the shard is the whole fleet
        return raw_scale, status, n_features
    max_len = max(e - s for s, e in shards)
    mine = pack_results(raw_scale, status, n_features, max_len)
    full = torch.empty(world, max_len, 3, dtype=torch.float64, device=mine.device)
    dist.all_gather_into_tensor(full.view(world * max_len, 3), mine, group=group)
    if all(e - s == max_len for s, e in shards):              # equal shards (weak scaling): no padding to strip
        cat = full.view(world * max_len, 3)
    else:
        cat = torch.cat([full[r, : e - s] for r, (s, e) in enumerate(shards)], 0)
    return cat[:, 0].contiguous(), cat[:, 1].to(torch.uint8).contiguous(), cat[:, 2].to(torch.int32).contiguous()
