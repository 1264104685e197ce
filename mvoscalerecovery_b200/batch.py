"""Batched host API: PyTorch tensors as the frame-batch container, libmvosr.so as the engine.

``ScaleRecovery`` owns one native handle (one CUDA device).  Inputs are CSR-packed frames
(structure-of-arrays float32 CUDA tensors + int32 offsets); every method enqueues the
CUDA kernels on torch's current stream and returns CUDA tensors without synchronising.

Reference path replaced: the per-frame loop of src/main_offline.py:57-88 around
rescale.ScaleEstimator.scale_calculation (src/rescale.py:113-178, 191-193) -- see
include/mvosr.h for the call-by-call mapping.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _native as N


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _chk(t: torch.Tensor, dtype, name: str, device, numel: Optional[int] = None):
    if not isinstance(t, torch.Tensor) or t.dtype != dtype or not t.is_contiguous() or t.device != device:
        raise ValueError("%s must be a contiguous %s tensor on %s (got %s, %s)" % (name, dtype, device, getattr(t, "dtype", type(t)), getattr(t, "device", None)))
    if numel is not None and t.numel() < numel:
        raise ValueError("%s holds %d elements, the call needs %d" % (name, t.numel(), numel))


def _chk_opt(t, dtype, name: str, device, numel: Optional[int] = None):
    """Optional device tensor: None passes, anything else is checked like a mandatory one (the C side dereferences it)."""
    if t is not None:
        _chk(t, dtype, name, device, numel)


def _host_array(a, dtype, name: str, numel: int):
    """A host buffer handed to the C ABI by raw pointer: numpy array or CPU torch tensor of exactly this dtype, C-contiguous,
    at least `numel` elements (the C side reinterprets the bytes: a float64 or strided view would be read as garbage)."""
    if isinstance(a, torch.Tensor):
        if a.device.type != "cpu" or a.dtype != getattr(torch, np.dtype(dtype).name) or not a.is_contiguous() or a.numel() < numel:
            raise ValueError("%s must be a contiguous CPU %s tensor with >= %d elements (got %s on %s, %d)"
                             % (name, np.dtype(dtype).name, numel, a.dtype, a.device, a.numel()))
        return C.c_void_p(a.data_ptr())
    if not isinstance(a, np.ndarray) or a.dtype != np.dtype(dtype) or not a.flags["C_CONTIGUOUS"] or a.size < numel:
        raise ValueError("%s must be a C-contiguous %s numpy array with >= %d elements (got %s)"
                         % (name, np.dtype(dtype).name, numel, getattr(a, "dtype", type(a))))
    return C.c_void_p(a.ctypes.data)


class ScaleRecovery:
    """Batch engine for stages 1-6.  ``config`` overrides fields of the reference defaults
    (``mvosr_default_config``), e.g. ``absolute_reference=1.7, ransac_iterations=256``."""

    def __init__(self, device: Optional[int] = None, **config):
        if not torch.cuda.is_available():
            raise RuntimeError("mvoscalerecovery_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.lib = N.lib()
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        cfg = N.default_config()
        if "grid_density" in config:            # tuning knob of the Delaunay stage: mean points per grid cell (reserved[0] = x100)
            cfg.reserved[0] = int(round(100 * float(config.pop("grid_density"))))
        if "window_factor" in config:           # tuning knob of the strip index: half-width of the candidate window in cell sides (reserved[1] = x100)
            cfg.reserved[1] = int(round(100 * float(config.pop("window_factor"))))
        for k, v in config.items():
            if not hasattr(cfg, k):
                raise TypeError("unknown config field %r" % k)
            setattr(cfg, k, v)
        self.config = cfg
        h = C.c_void_p()
        N.check(self.lib.mvosr_create(C.byref(cfg), self.device_index, C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self.lib.mvosr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launch_count(self) -> int:
        return int(self.lib.mvosr_launch_count(self._h))

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ------------------------------------------------------------------ stage 1
    def triangulate_frames(self, offsets, cur_u, cur_v, ref_u, ref_v, poses, e_mask=None):
        """Replaces cv2.recoverPose's triangulation + main.py:102-104. Returns dict(x,y,z,u,v,n_out)."""
        dev = self.device
        F = offsets.numel() - 1
        _chk(offsets, torch.int32, "offsets", dev)
        for n, t in (("cur_u", cur_u), ("cur_v", cur_v), ("ref_u", ref_u), ("ref_v", ref_v)):
            _chk(t, torch.float32, n, dev)
        _chk(poses, torch.float64, "poses", dev, 12 * F)
        _chk_opt(e_mask, torch.uint8, "e_mask", dev, cur_u.numel())
        M = cur_u.numel()
        out = {k: torch.empty(M, dtype=torch.float32, device=dev) for k in "xyzuv"}
        n_out = torch.zeros(F, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            N.check(self.lib.mvosr_triangulate_frames(self._h, F, _ptr(offsets), _ptr(cur_u), _ptr(cur_v), _ptr(ref_u), _ptr(ref_v),
                                                      _ptr(e_mask), _ptr(poses), _ptr(out["x"]), _ptr(out["y"]), _ptr(out["z"]),
                                                      _ptr(out["u"]), _ptr(out["v"]), _ptr(n_out), self._stream()))
        out["n_out"] = n_out
        return out

    def find_essential_frames(self, offsets, cur_u, cur_v, ref_u, ref_v, hypotheses: int = 1000, threshold: float = 0.5, seed: int = 0,
                              frame_index=None, seq_id: int = 0, confidence: float = 0.999):
        """Replaces cv2.findEssentialMat(px_cur, px_ref, K, RANSAC, 0.999, threshold) (visual_odometry.py:100-102,129-130) for F
        frames: five-point RANSAC on the Philox stream; hypotheses = OpenCV's maxIters, confidence = its prob (the adaptive count,
        checked every 128 hypotheses; 0 = run them all).  Returns dict(essential (F,9) f64, e_mask (M,) u8, n_inliers, best_hyp,
        hyps_used (F,))."""
        dev = self.device
        F = offsets.numel() - 1
        _chk(offsets, torch.int32, "offsets", dev)
        for n, t in (("cur_u", cur_u), ("cur_v", cur_v), ("ref_u", ref_u), ("ref_v", ref_v)):
            _chk(t, torch.float32, n, dev)
        if frame_index is not None:
            _chk(frame_index, torch.int32, "frame_index", dev)
        essential = torch.zeros((F, 9), dtype=torch.float64, device=dev)
        e_mask = torch.zeros(cur_u.numel(), dtype=torch.uint8, device=dev)
        n_inliers = torch.zeros(F, dtype=torch.int32, device=dev)
        best_hyp = torch.full((F,), -1, dtype=torch.int32, device=dev)
        hyps_used = torch.zeros(F, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            N.check(self.lib.mvosr_find_essential_frames(self._h, F, _ptr(offsets), _ptr(cur_u), _ptr(cur_v), _ptr(ref_u), _ptr(ref_v),
                                                         int(hypotheses), float(threshold), float(confidence), int(seed), _ptr(frame_index), int(seq_id),
                                                         _ptr(essential), _ptr(e_mask), _ptr(n_inliers), _ptr(best_hyp), _ptr(hyps_used), self._stream()))
        return dict(essential=essential, e_mask=e_mask, n_inliers=n_inliers, best_hyp=best_hyp, hyps_used=hyps_used)

    def recover_pose_frames(self, offsets, cur_u, cur_v, ref_u, ref_v, essential, e_mask=None):
        """The pose selection of cv2.recoverPose (visual_odometry.py:129-133): essential (F,9) float64 -> dict(poses (F,12), n_good (F,4))."""
        dev = self.device
        F = offsets.numel() - 1
        _chk(offsets, torch.int32, "offsets", dev)
        for n, t in (("cur_u", cur_u), ("cur_v", cur_v), ("ref_u", ref_u), ("ref_v", ref_v)):
            _chk(t, torch.float32, n, dev)
        _chk(essential, torch.float64, "essential", dev, 9 * F)
        _chk_opt(e_mask, torch.uint8, "e_mask", dev, cur_u.numel())
        poses = torch.empty((F, 12), dtype=torch.float64, device=dev)
        n_good = torch.zeros((F, 4), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            N.check(self.lib.mvosr_recover_pose_frames(self._h, F, _ptr(offsets), _ptr(cur_u), _ptr(cur_v), _ptr(ref_u), _ptr(ref_v), _ptr(e_mask),
                                                       _ptr(essential), _ptr(poses), _ptr(n_good), self._stream()))
        return dict(poses=poses, n_good=n_good)

    # ------------------------------------------------------------------ stages 2-5
    def scale_frames(self, offsets, x, y, z, u, v, max_features: int, counts=None, frame_index0: int = 0, seq_id: int = 0,
                     seed: int = 0, stats: bool = True, debug: bool = False):
        """Replaces feature_selection + RANSAC + height/scale of rescale.py:113-167 for F frames.
        Returns dict(raw_scale f64[F], status u8[F], stats (structured numpy view factory), debug buffers)."""
        dev = self.device
        F = offsets.numel() - 1
        _chk(offsets, torch.int32, "offsets", dev)
        for n, t in (("x", x), ("y", y), ("z", z), ("u", u), ("v", v)):
            _chk(t, torch.float32, n, dev)
        _chk_opt(counts, torch.int32, "counts", dev, F)
        M = x.numel()
        raw = torch.empty(F, dtype=torch.float64, device=dev)
        status = torch.zeros(F, dtype=torch.uint8, device=dev)
        st = torch.zeros(F * C.sizeof(N.FrameStats), dtype=torch.uint8, device=dev) if stats else None
        dbg_struct = None
        dbg = {}
        if debug:
            dbg = dict(tri1=torch.full((2 * M, 3), -1, dtype=torch.int32, device=dev), n_tri1=torch.zeros(F, dtype=torch.int32, device=dev),
                       keep=torch.zeros(M, dtype=torch.uint8, device=dev), tri2=torch.full((2 * M, 3), -1, dtype=torch.int32, device=dev),
                       tri_flags=torch.zeros(2 * M, dtype=torch.uint8, device=dev), tri_height=torch.zeros(2 * M, dtype=torch.float64, device=dev),
                       inlier=torch.zeros(M, dtype=torch.uint8, device=dev), data_id=torch.full((6 * M,), -1, dtype=torch.int32, device=dev))
            dbg_struct = N.DebugBuffers(*[dbg[k].data_ptr() for k in ("tri1", "n_tri1", "keep", "tri2", "tri_flags", "tri_height", "inlier", "data_id")])
        with torch.cuda.device(dev):
            N.check(self.lib.mvosr_scale_frames(self._h, F, _ptr(offsets), _ptr(counts), _ptr(x), _ptr(y), _ptr(z), _ptr(u), _ptr(v),
                                                int(max_features), int(frame_index0), int(seq_id), C.c_uint64(int(seed)),
                                                _ptr(raw), _ptr(status), _ptr(st), C.byref(dbg_struct) if dbg_struct else None,
                                                self._stream()))
        return dict(raw_scale=raw, status=status, stats=st, debug=dbg)

    def scale_frames_f64(self, offsets, feature3d, feature2d, max_features: int, frame_index0: int = 0, seq_id: int = 0,
                         seed: int = 0, stats: bool = True, debug: bool = False):
        """scale_frames on the float64 arrays of the reference's own hand-off: feature3d (M,3), feature2d (M,2) float64 CUDA tensors
        (array-of-structures, numpy's layout), CSR by offsets.  Gates, votes, planes and RANSAC evaluate the float64 values
        (mvosr_scale_frames_f64)."""
        dev = self.device
        F = offsets.numel() - 1
        _chk(offsets, torch.int32, "offsets", dev)
        M = feature3d.shape[0]
        _chk(feature3d, torch.float64, "feature3d", dev, 3 * M)
        _chk(feature2d, torch.float64, "feature2d", dev, 2 * M)
        raw = torch.empty(F, dtype=torch.float64, device=dev)
        status = torch.zeros(F, dtype=torch.uint8, device=dev)
        st = torch.zeros(F * C.sizeof(N.FrameStats), dtype=torch.uint8, device=dev) if stats else None
        dbg_struct, dbg = None, {}
        if debug:
            dbg = dict(tri1=torch.full((2 * M, 3), -1, dtype=torch.int32, device=dev), n_tri1=torch.zeros(F, dtype=torch.int32, device=dev),
                       keep=torch.zeros(M, dtype=torch.uint8, device=dev), tri2=torch.full((2 * M, 3), -1, dtype=torch.int32, device=dev),
                       tri_flags=torch.zeros(2 * M, dtype=torch.uint8, device=dev), tri_height=torch.zeros(2 * M, dtype=torch.float64, device=dev),
                       inlier=torch.zeros(M, dtype=torch.uint8, device=dev), data_id=torch.full((6 * M,), -1, dtype=torch.int32, device=dev))
            dbg_struct = N.DebugBuffers(*[dbg[k].data_ptr() for k in ("tri1", "n_tri1", "keep", "tri2", "tri_flags", "tri_height", "inlier", "data_id")])
        with torch.cuda.device(dev):
            N.check(self.lib.mvosr_scale_frames_f64(self._h, F, _ptr(offsets), _ptr(feature3d), _ptr(feature2d), int(max_features),
                                                    int(frame_index0), int(seq_id), C.c_uint64(int(seed)), _ptr(raw), _ptr(status), _ptr(st),
                                                    C.byref(dbg_struct) if dbg_struct else None, self._stream()))
        return dict(raw_scale=raw, status=status, stats=st, debug=dbg)

    def scale_frame_host_f64(self, feature3d: np.ndarray, feature2d: np.ndarray, frame_index: int = 0, seq_id: int = 0, seed: int = 0,
                             stats: bool = True):
        """One frame, numpy float64 (n,3) / (n,2) in, (raw_scale, status, n_features, stats or None) out: the whole per-frame call
        of the drop-in in one C-ABI call (mvosr_scale_frame_host_f64).  Synchronises."""
        n = int(feature3d.shape[0])
        p3 = _host_array(feature3d, np.float64, "feature3d", 3 * n)
        p2 = _host_array(feature2d, np.float64, "feature2d", 2 * n)
        if feature2d.shape[0] != n:
            raise ValueError("feature3d and feature2d must hold the same number of features")
        rec = N.FrameRecord()
        st = N.FrameStats() if stats else None
        N.check(self.lib.mvosr_scale_frame_host_f64(self._h, n, p3, p2, int(frame_index), int(seq_id), C.c_uint64(int(seed)),
                                                    C.byref(rec), C.byref(st) if stats else None))
        return rec.raw_scale, int(rec.status), int(rec.n_features), st

    def scale_frames_from_correspondences(self, offsets, cur_u, cur_v, ref_u, ref_v, poses, max_features: int, e_mask=None,
                                          frame_index0: int = 0, seq_id: int = 0, seed: int = 0, stats: bool = False):
        """Stages 1-5 fused (correspondences + poses -> raw scales)."""
        dev = self.device
        F = offsets.numel() - 1
        _chk(offsets, torch.int32, "offsets", dev)
        for n, t in (("cur_u", cur_u), ("cur_v", cur_v), ("ref_u", ref_u), ("ref_v", ref_v)):
            _chk(t, torch.float32, n, dev)
        _chk(poses, torch.float64, "poses", dev, 12 * F)
        _chk_opt(e_mask, torch.uint8, "e_mask", dev, cur_u.numel())
        raw = torch.empty(F, dtype=torch.float64, device=dev)
        status = torch.zeros(F, dtype=torch.uint8, device=dev)
        nfeat = torch.zeros(F, dtype=torch.int32, device=dev)
        st = torch.zeros(F * C.sizeof(N.FrameStats), dtype=torch.uint8, device=dev) if stats else None
        with torch.cuda.device(dev):
            N.check(self.lib.mvosr_scale_frames_from_correspondences(
                self._h, F, _ptr(offsets), _ptr(cur_u), _ptr(cur_v), _ptr(ref_u), _ptr(ref_v), _ptr(e_mask), _ptr(poses),
                int(max_features), int(frame_index0), int(seq_id), C.c_uint64(int(seed)), _ptr(raw), _ptr(status), _ptr(nfeat), _ptr(st),
                self._stream()))
        return dict(raw_scale=raw, status=status, n_features=nfeat, stats=st)

    def bucket_frames(self, offsets, u, v, bucket_size: int = 30, density: int = 2, seed: int = 0, frame_index=None, seq_id: int = 0):
        """bucket(features, bucket_size, density) of detector.py:65-95 for F frames: dict(index (M,) int32 = survivors' positions inside
        their frame, compacted at the frame's base offset, n_out (F,), status (F,))."""
        dev = self.device
        F = offsets.numel() - 1
        _chk(offsets, torch.int32, "offsets", dev)
        _chk(u, torch.float32, "u", dev)
        _chk(v, torch.float32, "v", dev)
        if frame_index is not None:
            _chk(frame_index, torch.int32, "frame_index", dev)
        index = torch.full((u.numel(),), -1, dtype=torch.int32, device=dev)
        n_out = torch.zeros(F, dtype=torch.int32, device=dev)
        status = torch.zeros(F, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            N.check(self.lib.mvosr_bucket_frames(self._h, F, _ptr(offsets), _ptr(u), _ptr(v), int(bucket_size), int(density), int(seed),
                                                 _ptr(frame_index), int(seq_id), _ptr(index), _ptr(n_out), _ptr(status), self._stream()))
        return dict(index=index, n_out=n_out, status=status)

    def pose_mask_frames(self, offsets, cur_u, cur_v, ref_u, ref_v, poses, e_mask=None):
        """recoverPose's per-correspondence mask under the given poses, ANDed with e_mask (visual_odometry.py:134-136): (M,) uint8,
        set exactly for the correspondences triangulate_frames keeps."""
        dev = self.device
        F = offsets.numel() - 1
        _chk(offsets, torch.int32, "offsets", dev)
        for n, t in (("cur_u", cur_u), ("cur_v", cur_v), ("ref_u", ref_u), ("ref_v", ref_v)):
            _chk(t, torch.float32, n, dev)
        _chk(poses, torch.float64, "poses", dev, 12 * F)
        _chk_opt(e_mask, torch.uint8, "e_mask", dev, cur_u.numel())
        mask = torch.zeros(cur_u.numel(), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            N.check(self.lib.mvosr_pose_mask_frames(self._h, F, _ptr(offsets), _ptr(cur_u), _ptr(cur_v), _ptr(ref_u), _ptr(ref_v), _ptr(e_mask),
                                                    _ptr(poses), _ptr(mask), self._stream()))
        return mask

    def scale_frames_from_tracks(self, offsets, cur_u, cur_v, ref_u, ref_v, max_features: int, hypotheses: int = 1000, threshold: float = 0.5,
                                 frame_index0: int = 0, seq_id: int = 0, seed: int = 0, stats: bool = False, confidence: float = 0.999):
        """Tracked correspondences alone -> raw scales: the geometry of VisualOdometry.processFrame (visual_odometry.py:129-147:
        findEssentialMat, recoverPose, triangulation) followed by the scale recovery of rescale.py:113-167, three kernels on one
        stream with no host round trip.  Returns scale_frames_from_correspondences' dict + essential, e_mask, n_inliers, poses."""
        F = offsets.numel() - 1
        fi = None if frame_index0 == 0 else torch.arange(frame_index0, frame_index0 + F, dtype=torch.int32, device=self.device)
        ess = self.find_essential_frames(offsets, cur_u, cur_v, ref_u, ref_v, hypotheses=hypotheses, threshold=threshold, seed=seed,
                                         frame_index=fi, seq_id=seq_id, confidence=confidence)
        pose = self.recover_pose_frames(offsets, cur_u, cur_v, ref_u, ref_v, ess["essential"])     # over all tracks, as the reference calls it
        out = self.scale_frames_from_correspondences(offsets, cur_u, cur_v, ref_u, ref_v, pose["poses"], max_features, e_mask=ess["e_mask"],
                                                     frame_index0=frame_index0, seq_id=seq_id, seed=seed, stats=stats)
        out.update(essential=ess["essential"], e_mask=ess["e_mask"], n_inliers=ess["n_inliers"], hyps_used=ess["hyps_used"], poses=pose["poses"])
        return out

    def scale_shard_from_correspondences(self, offsets, cur_u, cur_v, ref_u, ref_v, poses, max_features: int, records,
                                         frame_seq=None, frame_index=None, frame_index0: int = 0, seq_id: int = 0, order=None,
                                         e_mask=None, seed: int = 0):
        """Stages 1-5 over one rank's shard of a fleet in ONE launch (mvosr_scale_shard_from_correspondences): the frame range may
        span sequences -- ``frame_seq`` / ``frame_index`` (int32 [F]) carry every frame's Philox stream --, ``order`` (int32 [F]) is
        the processing order (largest frames first), and the results land as 16-byte records in ``records`` (uint8 [>= F, 16], e.g.
        this rank's block of the all-gather buffer, so the collective runs in place)."""
        dev = self.device
        F = offsets.numel() - 1
        _chk(offsets, torch.int32, "offsets", dev)
        for n, t in (("cur_u", cur_u), ("cur_v", cur_v), ("ref_u", ref_u), ("ref_v", ref_v)):
            _chk(t, torch.float32, n, dev)
        _chk(poses, torch.float64, "poses", dev, 12 * F)
        _chk(records, torch.uint8, "records", dev, N.RECORD_BYTES * F)
        _chk_opt(frame_seq, torch.int32, "frame_seq", dev, F)
        _chk_opt(frame_index, torch.int32, "frame_index", dev, F)
        _chk_opt(order, torch.int32, "order", dev, F)
        _chk_opt(e_mask, torch.uint8, "e_mask", dev, cur_u.numel())
        with torch.cuda.device(dev):
            N.check(self.lib.mvosr_scale_shard_from_correspondences(
                self._h, F, _ptr(offsets), _ptr(cur_u), _ptr(cur_v), _ptr(ref_u), _ptr(ref_v), _ptr(e_mask), _ptr(poses), int(max_features),
                _ptr(frame_seq), _ptr(frame_index), int(frame_index0), int(seq_id), _ptr(order), C.c_uint64(int(seed)), _ptr(records),
                self._stream()))
        return records

    def filter_records(self, seq_offsets, records, slot=None, move_flags=None, filter10: bool = True, out=None):
        """Stage 6 straight from gathered records (mvosr_filter_records): frame f of the global order is records[slot[f]]."""
        dev = self.device
        _chk(seq_offsets, torch.int32, "seq_offsets", dev)
        _chk(records, torch.uint8, "records", dev)
        S = seq_offsets.numel() - 1
        F = slot.numel() if slot is not None else records.numel() // N.RECORD_BYTES
        _chk_opt(slot, torch.int32, "slot", dev)
        _chk_opt(move_flags, torch.uint8, "move_flags", dev, F)
        if out is None:
            out = dict(scale=torch.empty(F, dtype=torch.float64, device=dev),
                       filter10=torch.empty(F, dtype=torch.float64, device=dev) if filter10 else None)
        with torch.cuda.device(dev):
            N.check(self.lib.mvosr_filter_records(self._h, S, _ptr(seq_offsets), _ptr(records), _ptr(slot), _ptr(move_flags),
                                                  _ptr(out["scale"]), _ptr(out.get("filter10")), self._stream()))
        return out

    # ------------------------------------------------------------------ stage 6
    def filter_sequences(self, seq_offsets, raw_scale, status, move_flags=None, n_features=None, filter10: bool = True):
        """Replaces the gating of main_offline.py:57-88 + rescale.py:168-178, then evaluate_scale.filter(...,10)."""
        dev = self.device
        _chk(seq_offsets, torch.int32, "seq_offsets", dev)
        _chk(raw_scale, torch.float64, "raw_scale", dev)
        _chk(status, torch.uint8, "status", dev)
        S = seq_offsets.numel() - 1
        F = raw_scale.numel()
        _chk(status, torch.uint8, "status", dev, F)
        _chk_opt(move_flags, torch.uint8, "move_flags", dev, F)
        _chk_opt(n_features, torch.int32, "n_features", dev, F)
        out = torch.empty(F, dtype=torch.float64, device=dev)
        f10 = torch.empty(F, dtype=torch.float64, device=dev) if filter10 else None
        with torch.cuda.device(dev):
            N.check(self.lib.mvosr_filter_sequences(self._h, S, _ptr(seq_offsets), _ptr(raw_scale), _ptr(status), _ptr(move_flags),
                                                    _ptr(n_features), _ptr(out), _ptr(f10), self._stream()))
        return dict(scale=out, filter10=f10)

    # ------------------------------------------------------------------ Delaunay alone
    def delaunay_frames(self, offsets, u, v, max_features: int):
        """Canonical triangles per frame (replacement of scipy.spatial.Delaunay(...).simplices)."""
        dev = self.device
        F = offsets.numel() - 1
        _chk(offsets, torch.int32, "offsets", dev)
        _chk(u, torch.float32, "u", dev)
        _chk(v, torch.float32, "v", dev)
        M = u.numel()
        tri = torch.full((2 * M, 3), -1, dtype=torch.int32, device=dev)
        ntri = torch.zeros(F, dtype=torch.int32, device=dev)
        status = torch.zeros(F, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            N.check(self.lib.mvosr_delaunay_frames(self._h, F, _ptr(offsets), _ptr(u), _ptr(v), int(max_features), _ptr(tri), _ptr(ntri),
                                                   _ptr(status), self._stream()))
        return dict(tri=tri, n_tri=ntri, status=status)

    # ------------------------------------------------------------------ stand-alone primitives (caller-supplied triangles / points)
    def triangle_planes(self, tri, xyz):
        """n = P^-1 1 per triangle (rescale.py:77-84).  tri int32 (T,3), xyz float64 (N,3) CUDA tensors.
        Returns dict(normal (T,3), height (T,), mean_y (T,))."""
        dev = self.device
        _chk(tri, torch.int32, "tri", dev)
        _chk(xyz, torch.float64, "xyz", dev)
        T = tri.shape[0]
        normal = torch.empty((T, 3), dtype=torch.float64, device=dev)
        height = torch.empty(T, dtype=torch.float64, device=dev)
        mean_y = torch.empty(T, dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            N.check(self.lib.mvosr_triangle_planes(self._h, T, _ptr(tri), _ptr(xyz), _ptr(normal), _ptr(height), _ptr(mean_y), self._stream()))
        return dict(normal=normal, height=height, mean_y=mean_y)

    def triangle_votes(self, tri, v, d):
        """Depth-order votes per vertex (rescale.py:45-72): dict(flagged int32 (N,), incident int32 (N,))."""
        dev = self.device
        _chk(tri, torch.int32, "tri", dev)
        _chk(v, torch.float64, "v", dev)
        _chk(d, torch.float64, "d", dev)
        n = v.numel()
        flagged = torch.empty(n, dtype=torch.int32, device=dev)
        incident = torch.empty(n, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            N.check(self.lib.mvosr_triangle_votes(self._h, tri.shape[0], _ptr(tri), _ptr(v), _ptr(d), n, _ptr(flagged), _ptr(incident), self._stream()))
        return dict(flagged=flagged, incident=incident)

    def ransac_planes(self, offsets, xyz, iterations: int = 100, threshold: float = 0.005, goal_fraction: float = 0.8,
                      stop_at_goal: bool = True, seed: int = 0, frame_index=None, seq_id: int = 0):
        """Batched get_pitch_ransac (estimate_road_norm.py:66-70) over S point lists (CSR offsets into xyz float64 (M,3)).
        Returns dict(model (S,4), ic, best_hyp, hyps_used)."""
        dev = self.device
        _chk(offsets, torch.int32, "offsets", dev)
        _chk(xyz, torch.float64, "xyz", dev)
        S = offsets.numel() - 1
        model = torch.empty((S, 4), dtype=torch.float64, device=dev)
        ic = torch.zeros(S, dtype=torch.int32, device=dev)
        best = torch.zeros(S, dtype=torch.int32, device=dev)
        used = torch.zeros(S, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            N.check(self.lib.mvosr_ransac_planes(self._h, S, _ptr(offsets), _ptr(xyz), int(iterations), float(threshold), float(goal_fraction),
                                                 int(bool(stop_at_goal)), C.c_uint64(int(seed)), _ptr(frame_index), int(seq_id),
                                                 _ptr(model), _ptr(ic), _ptr(best), _ptr(used), self._stream()))
        return dict(model=model, ic=ic, best_hyp=best, hyps_used=used)

    def integrate_paths(self, seq_offsets, motions, scales=None):
        """get_path / motion2pose (main_offline.py:95-119): (F,12) relative motions (+ per-frame scales) -> (F+S,12) poses,
        sequence s in rows seq_offsets[s]+s .. seq_offsets[s+1]+s (first row = identity)."""
        dev = self.device
        _chk(seq_offsets, torch.int32, "seq_offsets", dev)
        _chk(motions, torch.float64, "motions", dev)
        if scales is not None:
            _chk(scales, torch.float64, "scales", dev)
        S = seq_offsets.numel() - 1
        F = motions.shape[0]
        poses = torch.empty((F + S, 12), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            N.check(self.lib.mvosr_integrate_paths(self._h, S, _ptr(seq_offsets), _ptr(motions), _ptr(scales), _ptr(poses), self._stream()))
        return poses

    def depth_from_mesh(self, width: int, height: int, fx: float, fy: float, cx: float, cy: float, tri, uv, datas):
        """Reconstruct.depth_generate (reconstruct.py:91-107): dict(depth (H,W) float64, tri_id (H,W) int32)."""
        dev = self.device
        _chk(tri, torch.int32, "tri", dev)
        _chk(uv, torch.float64, "uv", dev)
        _chk(datas, torch.float64, "datas", dev)
        depth = torch.empty((height, width), dtype=torch.float64, device=dev)
        tri_id = torch.empty((height, width), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            N.check(self.lib.mvosr_depth_from_mesh(self._h, int(width), int(height), float(fx), float(fy), float(cx), float(cy),
                                                   tri.shape[0], _ptr(tri), _ptr(uv), _ptr(datas), _ptr(depth), _ptr(tri_id), self._stream()))
        return dict(depth=depth, tri_id=tri_id)

    # ------------------------------------------------------------------ host buffers end to end
    def recover_scales_host(self, offsets: np.ndarray, cur_u, cur_v, ref_u, ref_v, poses, move_flags=None, max_features: int = 0,
                            seq_id: int = 0, seed: int = 0, out=None, seq_offsets=None):
        """Host (ideally pinned) numpy/torch-CPU buffers in, filtered scales out: copies + stages 1-6 inside.
        ``seq_offsets`` (int32 (S+1,), host): the batch holds S sequences (mvosr_recover_fleet_host; sequence ids seq_id + s)."""
        if not isinstance(offsets, (np.ndarray, torch.Tensor)) or offsets.ndim != 1 or offsets.shape[0] < 1:
            raise ValueError("offsets must be a 1-D int32 array of F + 1 entries")
        F = int(offsets.shape[0] - 1)
        p_off = _host_array(offsets, np.int32, "offsets", F + 1)
        o = offsets.numpy() if isinstance(offsets, torch.Tensor) else offsets
        if F and (int(o[0]) != 0 or bool(np.any(np.diff(o) < 0))):
            raise ValueError("offsets must start at 0 and be non-decreasing")
        M = int(o[-1]) if F else 0
        p_cu, p_cv = _host_array(cur_u, np.float32, "cur_u", M), _host_array(cur_v, np.float32, "cur_v", M)
        p_ru, p_rv = _host_array(ref_u, np.float32, "ref_u", M), _host_array(ref_v, np.float32, "ref_v", M)
        p_pose = _host_array(poses, np.float64, "poses", 12 * F)
        p_move = None if move_flags is None else _host_array(move_flags, np.uint8, "move_flags", F)
        if out is None:
            out = dict(scale=np.empty(F, np.float64), raw_scale=np.empty(F, np.float64), status=np.empty(F, np.uint8))
        p_scale = _host_array(out["scale"], np.float64, "out['scale']", F)
        p_raw = None if out.get("raw_scale") is None else _host_array(out["raw_scale"], np.float64, "out['raw_scale']", F)
        p_st = None if out.get("status") is None else _host_array(out["status"], np.uint8, "out['status']", F)
        if not max_features:
            max_features = int(np.max(np.diff(o))) if F else 0
        with torch.cuda.device(self.device):
            if seq_offsets is None:
                N.check(self.lib.mvosr_recover_scales_host(self._h, F, p_off, p_cu, p_cv, p_ru, p_rv, p_pose,
                                                           p_move, int(max_features), int(seq_id), C.c_uint64(int(seed)),
                                                           p_scale, p_raw, p_st))
            else:
                so = np.ascontiguousarray(seq_offsets.numpy() if isinstance(seq_offsets, torch.Tensor) else seq_offsets, dtype=np.int32)
                if so.ndim != 1 or so.shape[0] < 1 or int(so[0]) != 0 or int(so[-1]) != F or bool(np.any(np.diff(so) < 0)):
                    raise ValueError("seq_offsets must run from 0 to the number of frames, non-decreasing")
                N.check(self.lib.mvosr_recover_fleet_host(self._h, int(so.shape[0] - 1), C.c_void_p(so.ctypes.data), p_off, p_cu, p_cv, p_ru, p_rv,
                                                          p_pose, p_move, int(max_features), int(seq_id), C.c_uint64(int(seed)),
                                                          p_scale, p_raw, p_st))
        return out


def stats_to_numpy(stats_tensor: torch.Tensor) -> np.ndarray:
    """View the per-frame mvosr_frame_stats bytes as a numpy structured array."""
    dt = np.dtype([(n, np.int32) for n in ("n_features", "n_roi", "n_dup", "n_kept", "n_tri", "n_loose", "n_tight", "n_valid",
                                           "best_hyp", "best_ic", "hyps_used", "n_degenerate", "n_deferred", "n_exact")]
                  + [("height_level", np.float64), ("model", np.float64, (4,)), ("height", np.float64)])
    assert dt.itemsize == C.sizeof(N.FrameStats)
    return stats_tensor.cpu().numpy().view(dt)


def pack_frames(f3_list, f2_list, device):
    """Pack per-frame (n,3)/(n,2) arrays into the CSR float32 structure-of-arrays batch."""
    sizes = [int(a.shape[0]) for a in f3_list]
    off = np.zeros(len(sizes) + 1, np.int32)
    np.cumsum(sizes, out=off[1:])
    f3 = np.concatenate([np.asarray(a, np.float32).reshape(-1, 3) for a in f3_list], 0) if sizes else np.zeros((0, 3), np.float32)
    f2 = np.concatenate([np.asarray(a, np.float32).reshape(-1, 2) for a in f2_list], 0) if sizes else np.zeros((0, 2), np.float32)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    return dict(offsets=t(off), x=t(f3[:, 0]), y=t(f3[:, 1]), z=t(f3[:, 2]), u=t(f2[:, 0]), v=t(f2[:, 1]),
                max_features=max(sizes) if sizes else 0)
