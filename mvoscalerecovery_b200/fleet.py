"""Multi-GPU plumbing: shard a fleet of sequences by frame range, gather raw scales, filter.

Frames are independent up to the temporal filter, so the data path has no
collective: every rank runs stages 1-5 on its contiguous frame range.  Only the per-frame raw
results (raw scale f64, status, feature count: 24 B/frame packed as 3 float64) are exchanged with
ONE all-gather; then the strictly sequential stage 6 (slew limiter + deque median,
src/rescale.py:168-178, and the gating of src/main_offline.py:57-88) runs on the full vector.
A 10-frame halo would be exact for the windowed medians but not for the slew limiter, whose
memory is unbounded; gather-then-filter is exact and the payload is a few hundred KB.

Two forms of the exchange:
  * ``gather_results``: three separate per-frame tensors packed, gathered, unpacked (any layout, any backend);
  * ``RecordExchange``: the shard's results are written by the frame kernel as 16-byte records straight into this rank's block
    of ONE buffer (``mvosr_scale_shard_from_correspondences``), a single in-place ``all_gather_into_tensor`` fills the other
    blocks, and the filter reads frame f at ``slot[f]`` of that buffer (``mvosr_filter_records``): no pack / cat / cast kernels
    between the frame kernel and the filter.  This is what ``bench.py --workload fleet`` times.

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


RECORD_BYTES = 16        # sizeof(mvosr_frame_record): raw_scale f64, n_features i32, status u8, 3 pad


def slot_map(shards: List[Tuple[int, int]]) -> np.ndarray:
    """Global frame index -> position in the rank-major gather buffer whose blocks are padded to the longest shard."""
    max_len = max(e - s for s, e in shards)
    slot = np.empty(shards[-1][1], dtype=np.int32)
    for r, (s, e) in enumerate(shards):
        slot[s:e] = r * max_len + np.arange(e - s, dtype=np.int32)
    return slot


def frame_tables(seq_starts: Sequence[int], lo: int, hi: int):
    """(sequence id, frame index inside its sequence) of the global frames lo..hi-1 of concatenated sequences with the given
    start offsets (len S+1): the Philox stream of a frame, wherever the shard boundaries fall."""
    starts = np.asarray(seq_starts, dtype=np.int64)
    g = np.arange(lo, hi, dtype=np.int64)
    sq = np.searchsorted(starts, g, side="right") - 1
    return sq.astype(np.int32), (g - starts[sq]).astype(np.int32)


class RecordExchange:
    """The fleet's one collective: every rank owns block ``rank`` of ``buffer`` (uint8 [world, max_len, 16]); ``mine`` is the
    view the frame kernel writes this rank's records into; ``gather()`` is one in-place all-gather."""

    def __init__(self, shards: List[Tuple[int, int]], rank: int, device, group=None):
        self.shards, self.rank, self.group = shards, rank, group
        self.world = len(shards)
        self.max_len = max(e - s for s, e in shards)
        self.buffer = torch.zeros(self.world, self.max_len, RECORD_BYTES, dtype=torch.uint8, device=device)
        self.mine = self.buffer[rank]
        self.slot = torch.from_numpy(slot_map(shards)).to(device)

    def gather(self):
        if self.world > 1:
            dist.all_gather_into_tensor(self.buffer.view(self.world * self.max_len, RECORD_BYTES), self.mine, group=self.group)
        return self.buffer


def records_to_numpy(records: torch.Tensor) -> np.ndarray:
    """View gathered records as a structured array (raw_scale, n_features, status)."""
    dt = np.dtype([("raw_scale", np.float64), ("n_features", np.int32), ("status", np.uint8), ("pad", np.uint8, (3,))])
    return records.detach().cpu().contiguous().numpy().reshape(-1, RECORD_BYTES).view(dt).reshape(-1)


class ShardRunner:
    """One rank's share of a fleet job, the loop of src/main_offline.py:57-88 over its frame range: the concatenated sequences
    (global frame order, ``seq_starts`` = their start offsets, len S+1) are cut into ``shards``; this rank holds the
    correspondences of frames ``shards[rank]`` as a CSR batch.  ``step()`` runs stages 1-5 on the shard (one launch; optionally
    the largest frames first -- measured on the B200: no gain on frames within +-5 % of each other, whose cost varies far more by
    content than by size), the one all-gather of 16-byte records, and stage 6 on the full vector: three launches, no host
    synchronisation.  ``step_host()`` is the same from HOST buffers: the shard is copied in chunks on a copy stream while the
    previous chunk computes, and the filtered scales are copied back.

    host: dict of CPU tensors / numpy arrays ``offsets`` int32 (F+1), ``cur_u, cur_v, ref_u, ref_v`` float32 (M),
    ``poses`` float64 (F,12) for this rank's frames; pinned tensors make ``step_host`` asynchronous."""

    def __init__(self, engine, host: dict, shards, rank: int, seq_starts, seed: int, max_features: int = 0, move_flags=None,
                 host_chunks: int = 4, group=None, filter10: bool = True, largest_first: bool = False):
        self.eng, self.rank, self.shards, self.seed = engine, rank, shards, int(seed)
        dev = engine.device
        as_t = lambda a: a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
        self.host = {k: as_t(host[k]) for k in ("offsets", "cur_u", "cur_v", "ref_u", "ref_v", "poses")}
        lo, hi = shards[rank]
        F = hi - lo
        off = self.host["offsets"].numpy()
        if off.shape[0] != F + 1:
            raise ValueError("this rank's batch holds %d frames, its shard %d" % (off.shape[0] - 1, F))
        sizes = np.diff(off)
        self.n_frames, self.total_frames = F, int(shards[-1][1])
        self.max_features = int(max_features) or (int(sizes.max()) if F else 0)
        self.dev = {k: torch.empty_like(v, device=dev) for k, v in self.host.items()}
        for k, v in self.host.items():
            self.dev[k].copy_(v, non_blocking=True)
        fseq, fidx = frame_tables(seq_starts, lo, hi)
        self.frame_seq, self.frame_index = torch.from_numpy(fseq).to(dev), torch.from_numpy(fidx).to(dev)
        # largest frames first: the persistent CTAs take frames in this order, so the last wave is made of the cheapest ones
        self.order = torch.from_numpy(np.argsort(-sizes, kind="stable").astype(np.int32)).to(dev) if largest_first else None
        self.ex = RecordExchange(shards, rank, dev, group)
        self.slot = None if all(e - s == self.ex.max_len for s, e in shards) else self.ex.slot      # equal shards: identity
        self.seq_off = torch.from_numpy(np.asarray(seq_starts, dtype=np.int32)).to(dev)
        self.move = None if move_flags is None else as_t(move_flags).to(dev)
        self.out = dict(scale=torch.empty(self.total_frames, dtype=torch.float64, device=dev),
                        filter10=torch.empty(self.total_frames, dtype=torch.float64, device=dev) if filter10 else None)
        # host pipeline: chunk boundaries in frames, per-chunk processing order, pinned result buffer
        nc = max(1, min(int(host_chunks), F // 256 if F >= 256 else 1))
        self.chunks = [(F * c // nc, F * (c + 1) // nc) for c in range(nc)]
        self.chunk_order = [torch.from_numpy(np.argsort(-sizes[a:b], kind="stable").astype(np.int32)).to(dev) if largest_first else None
                            for a, b in self.chunks]
        self.chunk_m = [(int(off[a]), int(off[b])) for a, b in self.chunks]
        self.copy_stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        self.comp_streams = [torch.cuda.Stream(device=dev) for _ in range(2)] if dev.type == "cuda" else None
        self.host_out = torch.empty(self.total_frames, dtype=torch.float64).pin_memory() if dev.type == "cuda" else None
        self.launches_per_step = 2 + 1      # shard kernel, filter kernel (+ the collective's own kernel when world > 1)

    def _launch(self, a: int, b: int, order):
        d, off = self.dev, self.dev["offsets"]
        self.eng.scale_shard_from_correspondences(
            off[a:b + 1], d["cur_u"], d["cur_v"], d["ref_u"], d["ref_v"], d["poses"][a:b], self.max_features, self.ex.mine[a:b],
            frame_seq=self.frame_seq[a:b], frame_index=self.frame_index[a:b], order=order, seed=self.seed)

    def finish(self):
        self.ex.gather()
        return self.eng.filter_records(self.seq_off, self.ex.buffer, self.slot, self.move, out=self.out)

    def step(self, events=None):
        """events: optional (e_start, e_kernel, e_gather, e_filter) CUDA events recorded around the three phases."""
        if events:
            events[0].record()
        self._launch(0, self.n_frames, self.order)
        if events:
            events[1].record()
        self.ex.gather()
        if events:
            events[2].record()
        out = self.eng.filter_records(self.seq_off, self.ex.buffer, self.slot, self.move, out=self.out)
        if events:
            events[3].record()
        return out

    def step_host(self):
        """Host buffers in, filtered scales of the whole fleet out (pinned host tensor); synchronises before returning."""
        cur = torch.cuda.current_stream(self.eng.device)
        cs = self.copy_stream
        cs.wait_stream(cur)                                  # the previous step's kernels have consumed the device buffers
        with torch.cuda.stream(cs):
            self.dev["offsets"].copy_(self.host["offsets"], non_blocking=True)
            self.dev["poses"].copy_(self.host["poses"], non_blocking=True)
        evs = []
        for m0, m1 in self.chunk_m:
            with torch.cuda.stream(cs):
                for k in ("cur_u", "cur_v", "ref_u", "ref_v"):
                    self.dev[k][m0:m1].copy_(self.host[k][m0:m1], non_blocking=True)
                ev = torch.cuda.Event(); ev.record(cs); evs.append(ev)
        # consecutive chunks on two compute streams: the tail of one launch overlaps the head of the next
        for st in self.comp_streams:
            st.wait_stream(cur)
        for i, ((a, b), order, ev) in enumerate(zip(self.chunks, self.chunk_order, evs)):
            st = self.comp_streams[i & 1]
            st.wait_event(ev)
            with torch.cuda.stream(st):
                self._launch(a, b, order)
        for st in self.comp_streams:
            cur.wait_stream(st)
        out = self.finish()
        self.host_out.copy_(out["scale"], non_blocking=True)
        cur.synchronize()
        return self.host_out

    @property
    def h2d_bytes(self):
        return sum(v.numel() * v.element_size() for v in self.host.values())

    @property
    def d2h_bytes(self):
        return 8 * self.total_frames
