#!/usr/bin/env python
"""bench.py -- frames/sec of full scale recovery (stages 1-6) on synthetic KITTI-shaped data.

Contract (driver-facing):
    python bench.py --gpus N --steps K --warmup W              # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path (oracle port)
Under torchrun (N>1) one rank per GPU; rank 0 prints ONE JSON line.

Workload (BASELINE.json configs[1]): one offline sequence of 4541 KITTI-00-shaped frames per GPU,
~2.5k tracked correspondences (~2k road-ROI features) per frame, 1241x376 camera, 1.7 m camera
height, 10 % outliers.  A "step" is one pass of the hot path over the whole batch: tracked
correspondences + relative poses in, filtered per-frame scales out.  At N>1 the fleet is N such
sequences, frames sharded by range (weak scaling), one all-gather of the raw per-frame results, then
the temporal filter on the full vector.

value   : frames/s with the batch already resident in HBM (device-timed with CUDA events).
e2e     : frames/s through the host-buffer C-ABI call (pinned host inputs, H2D + D2H inside the timed region); also the same
          call on pageable numpy arrays (`pageable_value`) and the per-frame drop-in class (`dropin_fps`).
roofline: the fused frame kernel against measured HBM bandwidth; algorithmic bytes per frame = 16*n + 64
          (SURVEY.md section 8d).  The path is instruction-issue bound by design: `issue_frac` and
          `winstr_per_frame` (warp-instructions from the committed ncu capture of the SAME kernel sources, checked by hash)
          say how close to the issue peak it runs.
fleet   : BASELINE configs[3] in the same invocation -- the 23 201 frames of 11 KITTI-shaped sequences cut into N frame ranges
          (strong scaling): one shard launch, one in-place all-gather of 16-byte records, one filter launch.  fleet.value at N
          divided by fleet.value at 1 is the strong-scaling factor.
cpu_baseline / --impl reference: the UNMODIFIED reference (oracle/_ref, staged by `make -C oracle ref`) when present, else the
          oracle port (kind says which).
"""
from __future__ import annotations

import argparse
import contextlib
import hashlib
import io
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_FRAMES = 4541
N_CORR = 2500
SEED = 20261017
METRIC = "frames/sec scale recovery"
KITTI_LENGTHS = (4541, 1101, 4661, 801, 271, 2761, 1101, 1101, 4071, 1591, 1201)      # sequences 00-10: 23 201 frames
WORKLOADS = {
    # name: (frames per GPU (None: the fleet), correspondences per frame, description)
    "kitti00": (N_FRAMES, N_CORR, "offline sequence per GPU, KITTI-00-shaped synthetic correspondences, ~2k road features/frame (BASELINE configs[1])"),
    "dense": (592, 25000, "dense-flow stress, ~20k road features/frame, large-frame mode (BASELINE configs[2]; 592 of the 4541 frames per GPU)"),
    "fleet": (None, N_CORR, "fleet batch: 11 KITTI 00-10-shaped sequences, 23 201 frames sharded by frame range (BASELINE configs[3])"),
    "kitti00-ground": (N_FRAMES, N_CORR, "as kitti00 with the features sampled on the GROUND (X in U(-8,8) m, Z in U(5,40) m, SURVEY 8d): image density ~ 1/(v-cy)^3"),
    "kitti00-clustered": (N_FRAMES, N_CORR, "as kitti00 with 70 % of the road features in 12 Gaussian clusters (textured patches) over a uniform background"),
}
WORKLOAD_KW = {"kitti00-ground": {"density": "ground"}, "kitti00-clustered": {"density": "clustered"}}       # synth.make_frame options


def make_workload(n_frames, n_corr, seq, **kw):
    from mvoscalerecovery_b200 import synth
    return synth.make_sequence(seed=SEED, n_frames=n_frames, n_corr=n_corr, seq=seq, outlier_frac=0.10, **kw)


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md recipe).  Sampled through NVML every 5 ms
    from a thread (the region is a few hundred ms: `nvidia-smi -lms` would deliver two or three lines at best); nvidia-smi
    is the fallback when the NVML binding is missing."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, gpu_index, uuid=None):
        self.gpu, self.uuid = gpu_index, uuid
        self.proc, self.nvml, self.handle = None, None, None
        self.lines, self.sm, self.mask = [], [], 0
        self.stop_flag = threading.Event()

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            if self.uuid:
                try:
                    h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(self.uuid)) if not str(self.uuid).startswith("GPU-") else str(self.uuid))
                except Exception:
                    h = None
            self.handle = h or pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            self.nvml = pynvml
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                try:
                    self.mask |= int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                except Exception:
                    self.mask |= int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
            except Exception:
                pass
            time.sleep(0.005)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag.set()
            self.t.join(timeout=1)
            reasons = sorted(name for bit, name in self.REASONS if self.mask & bit)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.smax, "reasons": reasons,
                    "samples": len(self.sm), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); smax.append(float(p[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


# ------------------------------------------------------------------------------------------ CPU arms
REF_STAGED = os.path.join(ROOT, "oracle", "_ref", "src")        # `make -C oracle ref`: the unmodified reference, staged for the box


def reference_src():
    for p in (os.environ.get("MVOSR_REFERENCE_SRC"), REF_STAGED):
        if p and os.path.isfile(os.path.join(p, "rescale.py")):
            return p
    return None


def _load_unmodified_reference(src):
    """The reference's own src/rescale.py, imported as it lies in `src`.  Two things of this image are patched around it, none
    inside it: matplotlib is not installed (imported at rescale.py:17, never used on this path) and numpy >= 1.24 dropped
    np.float (rescale.py:76)."""
    import types
    for m in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if not hasattr(np, "float"):
        np.float = float
    if src not in sys.path:
        sys.path.insert(0, src)
    with contextlib.redirect_stdout(io.StringIO()):
        import rescale
        import param
    return rescale, param


def _reference_frames(batch, frames, cam, src):
    """The reference's per-frame path on the given frames (single process): its OpenCV stage-1 call
    (cv2.recoverPose(E, cur, ref, K, distanceThresh=100), visual_odometry.py:132-147, E from the known pose), the reprojection of
    main.py:102-104, the n > 100 gate and rescale.ScaleEstimator.scale_calculation (main_offline.py:73-75) -- UNMODIFIED code,
    its prints swallowed (they dominate otherwise)."""
    import cv2
    rescale, param = _load_unmodified_reference(src)
    est = rescale.ScaleEstimator(absolute_reference=1.7, window_size=5)
    K = np.array([[cam.fx, 0, cam.cx], [0, cam.fy, cam.cy], [0, 0, 1.0]])
    n = 0
    sink = io.StringIO()
    for f in frames:
        a, e = batch.offsets[f], batch.offsets[f + 1]
        if e - a == 0:
            continue
        cur = np.stack([batch.cur_u[a:e], batch.cur_v[a:e]], 1)
        ref = np.stack([batch.ref_u[a:e], batch.ref_v[a:e]], 1)
        P = batch.poses[f].reshape(3, 4)
        t = P[:, 3]
        E = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]]) @ P[:, :3]
        _, _, _, mask, pts4 = cv2.recoverPose(E, cur, ref, cameraMatrix=K, distanceThresh=100)
        X = (pts4[:3] / pts4[3:4]).T[np.array(mask > 0).reshape(-1)]
        f2 = X[:, 0:2].copy()
        f2[:, 0] = f2[:, 0] * cam.fx / X[:, 2] + cam.cx
        f2[:, 1] = f2[:, 1] * cam.fx / X[:, 2] + cam.cy
        if X.shape[0] > param.minimum_feature_for_scale:
            with contextlib.redirect_stdout(sink):
                est.initial_estimation(P[:, 3])
                est.scale_calculation(X, f2)
            sink.seek(0); sink.truncate(0)
        n += 1
    return n


def _oracle_frames(batch, frames, cam):
    """Oracle port of stages 1-5 on the given frames (single process). Returns frames processed."""
    from oracle import pipeline as P
    n = 0
    for f in frames:
        a, e = batch.offsets[f], batch.offsets[f + 1]
        if e - a == 0:
            continue
        cur = np.stack([batch.cur_u[a:e], batch.cur_v[a:e]], 1)
        ref = np.stack([batch.ref_u[a:e], batch.ref_v[a:e]], 1)
        Pm = batch.poses[f].reshape(3, 4)
        X, m = P.triangulate_dlt(cur, ref, Pm[:, :3], Pm[:, 3], cam.fx, cam.fy, cam.cx, cam.cy)
        X = X[m]
        uv = P.reproject(X, cam.fx, cam.cx, cam.cy)
        f3 = X.astype(np.float32).astype(np.float64)
        f2 = uv.astype(np.float32).astype(np.float64)
        if f3.shape[0] > P.MIN_FEATURES:
            P.frame_raw_scale(f3, f2, SEED, f, 0, absolute_reference=1.7)
        n += 1
    return n


def _worker(args):
    seq, lo, hi, n_frames, n_corr, src = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from mvoscalerecovery_b200 import synth
    b = _WORK_CACHE.get((n_frames, n_corr, seq))
    if b is None:
        b = make_workload(n_frames, n_corr, seq)
        _WORK_CACHE[(n_frames, n_corr, seq)] = b
    t0 = time.perf_counter()
    n = _reference_frames(b, range(lo, hi), synth.Camera(), src) if src else _oracle_frames(b, range(lo, hi), synth.Camera())
    return n, time.perf_counter() - t0


_WORK_CACHE = {}


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores -- the UNMODIFIED reference from
    oracle/_ref when it was staged (kind "reference"), else the oracle port (kind "port") -- a bounded sample of the same workload
    per step: `per` frames on each of `workers` processes."""
    import multiprocessing as mp
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    src = None if args.ref_port else reference_src()
    cores = os.cpu_count() or 1
    workers = args.ref_cores or max(1, min(cores, 64))
    per = args.ref_frames or (2 if src else 6)         # frames per worker per step: the unmodified reference takes ~0.5 s per frame
    n_frames = workers * per
    ctx = mp.get_context("fork")
    n_corr = args.features or WORKLOADS[args.workload][1]
    _WORK_CACHE[(n_frames, n_corr, 0)] = make_workload(n_frames, n_corr, 0, **WORKLOAD_KW.get(args.workload, {}))      # generated once in the parent: forked workers share it
    jobs = [(0, w * per, (w + 1) * per, n_frames, n_corr, src) for w in range(workers)]
    with ctx.Pool(workers) as pool:
        for _ in range(args.warmup):
            pool.map(_worker, jobs)
        t0 = time.perf_counter()
        done, busy = 0, 0.0
        for _ in range(args.steps):
            res = pool.map(_worker, jobs)
            done += sum(n for n, _ in res); busy += sum(t for _, t in res)
        dt = time.perf_counter() - t0
    val = done / dt
    kind = "reference" if src else "port"
    what = ("unmodified reference (oracle/_ref/src: cv2.recoverPose + rescale.ScaleEstimator.scale_calculation, main_offline.py:57-88)"
            if src else "oracle/pipeline.py stages 1-5 (vectorised numpy port; the reference was not staged: make -C oracle ref)")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload][2], "frames_per_step": n_frames, "correspondences_per_frame": n_corr,
                       "note": "CPU arm: a bounded sample of the workload per step; it does not scale with --gpus"},
            "cpu_baseline": {"value": val, "unit": "frames/s", "cores": workers, "kind": kind,
                             "frames_per_s_per_core": done / busy if busy > 0 else None,
                             "sample": "%d frames/step (%d per worker process), %s" % (n_frames, per, what)},
            "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return line


def cpu_baseline_single(n_sample, workload, features):
    """One host core on the first `n_sample` frames of the workload, in a child process (the reference's module names --
    rescale, graph, param ... -- must not meet the drop-in's in one interpreter): the reference arm with one worker."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--ref-cores", "1", "--ref-frames", str(n_sample),
           "--steps", "1", "--warmup", "0", "--workload", workload]
    if features:
        cmd += ["--features", str(features)]
    env = dict(os.environ, RANK="0", WORLD_SIZE="1", OMP_NUM_THREADS="1")
    try:
        p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900, env=env)
        line = json.loads([ln for ln in p.stdout.splitlines() if ln.startswith("{")][-1])
        cb = line["cpu_baseline"]
        cb["cores"] = 1
        cb["sample"] = "first %d frames of the workload on one core, %.1f s; %s" % (n_sample, line["ms_per_step"] / 1e3, cb["sample"].split(", ", 1)[1])
        return cb
    except Exception as ex:                      # the GPU line must not die with the CPU leg
        return {"value": None, "unit": "frames/s", "cores": 1, "kind": "unavailable", "sample": "cpu leg failed: %r" % (ex,)}


def build_rank_workload(workload, rank, world, frames=0, features=0, lengths=KITTI_LENGTHS, weighted=False):
    """What rank `rank` of `world` processes (host-side, no GPU): its CSR batch, the sequence pieces it consists of
    [(sequence, first frame, end frame)], every rank's frame range in the global frame order, and the sequence offsets of
    the whole job.  kitti00 / dense: weak scaling, sequence `rank` of the fleet on this rank.  fleet: strong scaling, the
    concatenated sequences cut into `world` contiguous frame ranges -- a range may span sequences; `weighted` balances the
    ranges by correspondences instead of frames (fleet.frame_shards(weights=...))."""
    from mvoscalerecovery_b200 import fleet, synth
    wl_frames, wl_corr, wl_desc = WORKLOADS[workload]
    n_corr = features or wl_corr
    kw = WORKLOAD_KW.get(workload, {})
    if workload == "fleet":
        seq_starts = np.concatenate([[0], np.cumsum(lengths)])
        total_frames = int(seq_starts[-1])
        weights = None
        if weighted:
            weights = np.concatenate([synth.sequence_sizes(SEED, L, n_corr, sq) for sq, L in enumerate(lengths)])
        shards = fleet.frame_shards(total_frames, world, weights)
        lo, hi = shards[rank]
        pieces = []
        for sq in range(len(lengths)):
            a, b = max(lo, int(seq_starts[sq])), min(hi, int(seq_starts[sq + 1]))
            if a < b:
                pieces.append((sq, a - int(seq_starts[sq]), b - int(seq_starts[sq])))
        parts = [synth.make_sequence(seed=SEED, n_frames=lengths[sq], n_corr=n_corr, seq=sq, outlier_frac=0.10, frame_range=(a, b))
                 for sq, a, b in pieces]
        n_frames = hi - lo
        seq_off_host = seq_starts.astype(np.int32)
    else:
        n_frames = frames or wl_frames
        total_frames = world * n_frames
        shards = [(r * n_frames, (r + 1) * n_frames) for r in range(world)]
        pieces = [(rank, 0, n_frames)]
        parts = [synth.make_sequence(seed=SEED, n_frames=n_frames, n_corr=n_corr, seq=rank, outlier_frac=0.10, **kw)]
        seq_off_host = np.arange(0, (world + 1) * n_frames, n_frames, dtype=np.int32)
    batch = parts[0] if len(parts) == 1 else synth.CorrespondenceBatch(
        np.concatenate([[0]] + [p.offsets[1:].astype(np.int64) + sum(int(q.offsets[-1]) for q in parts[:i]) for i, p in enumerate(parts)]).astype(np.int32),
        *[np.concatenate([getattr(p, k) for p in parts]) for k in ("cur_u", "cur_v", "ref_u", "ref_v", "poses", "move_flags", "true_scale")])
    return dict(batch=batch, pieces=pieces, shards=shards, seq_off_host=seq_off_host, total_frames=total_frames, n_frames=n_frames,
                n_corr=n_corr, desc=wl_desc)


# ------------------------------------------------------------------------------------------ GPU arm
STATUS_BITS = (("updated", 1), ("second_dt", 2), ("few_roi", 4), ("no_model", 8), ("bad_input", 16), ("overflow", 32), ("skipped", 64))


def kernel_source_hash():
    """sha256 over the CUDA sources of libmvosr.so (csrc/*.cu, *.cuh, *.h + include/mvosr.h): ties a committed ncu capture to the
    kernels it was taken from."""
    h = hashlib.sha256()
    csrc = os.path.join(ROOT, "mvoscalerecovery_b200", "csrc")
    files = sorted(os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cu", ".cuh", ".h")))
    for p in files + [os.path.join(ROOT, "include", "mvosr.h")]:
        h.update(os.path.basename(p).encode()); h.update(open(p, "rb").read())
    return h.hexdigest()[:16]


def timed_leg(runner, steps, warmup, world, dev, sample_clocks=None):
    """W warm-up steps, then K steps bracketed by barrier + synchronize; CUDA events on the launching stream around the whole
    region and around every phase of every step.  Returns the max-over-ranks milliseconds."""
    import torch
    import torch.distributed as dist
    ev = lambda: torch.cuda.Event(enable_timing=True)
    evs = [(ev(), ev(), ev(), ev()) for _ in range(steps)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(warmup, 3)):
        runner.step()
    barrier()
    if sample_clocks:
        sample_clocks.start()
    launches0 = runner.eng.launch_count
    e0, e1 = ev(), ev()
    barrier()
    e0.record()
    for i in range(steps):
        out = runner.step(evs[i])
    e1.record()
    barrier()
    clocks = sample_clocks.stop() if sample_clocks else None
    ms = [e0.elapsed_time(e1) / steps, float(np.mean([a.elapsed_time(b) for a, b, _, _ in evs])),
          float(np.mean([b.elapsed_time(c) for _, b, c, _ in evs])), float(np.mean([c.elapsed_time(d) for _, _, c, d in evs]))]
    t = torch.tensor(ms, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step, k_ms, g_ms, f_ms = (float(x) for x in t.tolist())
    return dict(ms_step=ms_step, kernel_ms=k_ms, gather_ms=g_ms, filter_ms=f_ms, launches=runner.eng.launch_count - launches0,
                clocks=clocks, out=out)


def timed_host_leg(fn, steps, world, dev):
    """Wall clock around `steps` calls of the host-buffer entry (each synchronises itself); max over ranks, seconds per step."""
    import torch
    import torch.distributed as dist
    for _ in range(2):
        fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    torch.cuda.synchronize()
    te = torch.tensor([(time.perf_counter() - t0) / steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    return float(te.item())


def make_runner(eng, wl, rank, dev):
    import torch
    from mvoscalerecovery_b200 import fleet
    b = wl["batch"]
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    host = dict(offsets=pin(b.offsets), cur_u=pin(b.cur_u), cur_v=pin(b.cur_v), ref_u=pin(b.ref_u), ref_v=pin(b.ref_v), poses=pin(b.poses))
    r = fleet.ShardRunner(eng, host, wl["shards"], rank, wl["seq_off_host"], SEED, largest_first=os.environ.get("MVOSR_BENCH_ORDER", "natural") == "lpt")
    mode = os.environ.get("MVOSR_BENCH_ORDER", "natural")          # experiments on the processing order
    if mode in ("random", "spt", "strided"):
        sizes = np.diff(b.offsets)
        if mode == "random":
            o = np.random.default_rng(1).permutation(sizes.shape[0])
        elif mode == "spt":
            o = np.argsort(sizes, kind="stable")
        else:                                                  # sorted by size, then dealt out so that one wave of 148 CTAs sees every size class
            srt = np.argsort(-sizes, kind="stable")
            o = np.concatenate([srt[k::31] for k in range(31)])
        r.order = torch.from_numpy(o.astype(np.int32)).to(dev)
    torch.cuda.synchronize()
    return r


def status_hist(runner):
    from mvoscalerecovery_b200 import fleet
    rec = fleet.records_to_numpy(runner.ex.buffer)
    st = rec["status"] if runner.slot is None else rec["status"][runner.ex.slot.cpu().numpy()]
    h = {name: int(np.count_nonzero(st & bit)) for name, bit in STATUS_BITS}
    h["held"] = int(np.count_nonzero((st & 1) == 0))
    h["frames"] = int(st.shape[0])
    return h


def dropin_leg(eng, batch, n_frames, dev):
    """Calls/s of the per-frame drop-in itself: the loop of src/main_offline.py:57-88 over `n_frames` frames through
    compat/rescale.ScaleEstimator (what the unmodified mains call), numpy float64 arrays in, Python floats out."""
    import torch
    compat = os.path.join(ROOT, "mvoscalerecovery_b200", "compat")
    if compat not in sys.path:
        sys.path.insert(0, compat)
    import rescale                                            # the drop-in (compat/rescale.py), as main_offline.py imports it
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    n_frames = min(n_frames, batch.n_frames)
    off = batch.offsets[: n_frames + 1]
    M = int(off[-1])
    s1 = eng.triangulate_frames(t(off), t(batch.cur_u[:M]), t(batch.cur_v[:M]), t(batch.ref_u[:M]), t(batch.ref_v[:M]), t(batch.poses[:n_frames]))
    n_out = s1["n_out"].cpu().numpy()
    xyz = np.stack([s1[k].cpu().numpy() for k in "xyz"], 1).astype(np.float64)
    uv = np.stack([s1[k].cpu().numpy() for k in "uv"], 1).astype(np.float64)
    f3s = [xyz[off[f]: off[f] + n_out[f]] for f in range(n_frames)]
    f2s = [uv[off[f]: off[f] + n_out[f]] for f in range(n_frames)]
    est = rescale.ScaleEstimator(absolute_reference=1.7, window_size=5)
    for f in range(min(20, n_frames)):
        est.scale_calculation(f3s[f], f2s[f])
    est = rescale.ScaleEstimator(absolute_reference=1.7, window_size=5)
    scales = [0]
    t0 = time.perf_counter()
    for f in range(n_frames):                                 # main_offline.py:57-88
        if f3s[f].shape[0] > 100:
            est.initial_estimation(batch.poses[f][3:12:4])
            scale, _ = est.scale_calculation(f3s[f], f2s[f])
            scales.append(scale)
        else:
            scales.append(scales[-1])
    dt = time.perf_counter() - t0
    return {"value": n_frames / dt, "unit": "frames/s", "frames": n_frames,
            "api": "compat/rescale.ScaleEstimator.scale_calculation per frame (numpy float64 in, one launch + one D2H per call)"}


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from mvoscalerecovery_b200.batch import ScaleRecovery

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE JSON line: whatever libraries print there (e.g. NCCL's version banner) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    eng = ScaleRecovery(device=local_rank, absolute_reference=1.7)
    try:
        uuid = torch.cuda.get_device_properties(local_rank).uuid
    except Exception:
        uuid = None

    # ---- the bench line's own workload
    wl = build_rank_workload(args.workload, rank, world, args.frames, args.features, weighted=True)
    batch, n_frames, total_frames, n_corr = wl["batch"], wl["n_frames"], wl["total_frames"], wl["n_corr"]
    runner = make_runner(eng, wl, rank, dev)
    main = timed_leg(runner, args.steps, args.warmup, world, dev, ClockSampler(local_rank, uuid))
    value = total_frames / (main["ms_step"] * 1e-3)
    hist = status_hist(runner)
    scales = main["out"]["scale"].cpu().numpy()

    # ---- end to end: host buffers in, filtered scales out, copies inside the timed region
    e2e_steps = max(1, min(args.steps, 5))
    host_kw = dict(max_features=runner.max_features, seq_id=wl["pieces"][0][0], seed=SEED)
    h = runner.host
    move_pin = torch.from_numpy(np.ascontiguousarray(batch.move_flags)).pin_memory()
    if args.workload == "fleet":
        e2e_s = timed_host_leg(runner.step_host, e2e_steps, world, dev)
        e2e_api = "fleet.ShardRunner.step_host (pinned host buffers -> chunked H2D -> shard kernel -> all-gather -> filter -> D2H)"
        h2d, d2h = runner.h2d_bytes, runner.d2h_bytes
        page_s = None
    else:
        res = dict(scale=np.empty(n_frames, np.float64), raw_scale=np.empty(n_frames, np.float64), status=np.empty(n_frames, np.uint8))
        call = lambda: eng.recover_scales_host(h["offsets"], h["cur_u"], h["cur_v"], h["ref_u"], h["ref_v"], h["poses"], move_pin, out=res, **host_kw)
        e2e_s = timed_host_leg(call, e2e_steps, world, dev)
        e2e_api = "mvosr_recover_scales_host (pinned host buffers, copies inside)"
        M = int(batch.offsets[-1])
        h2d = 4 * (n_frames + 1) + 16 * M + 96 * n_frames + n_frames + 8
        d2h = 8 * n_frames + 8 * n_frames + n_frames
        # the same call on pageable numpy arrays (what np.load of the reference's hand-off gives a caller)
        pg = lambda: eng.recover_scales_host(batch.offsets, batch.cur_u, batch.cur_v, batch.ref_u, batch.ref_v, batch.poses, batch.move_flags, out=res, **host_kw)
        page_s = timed_host_leg(pg, max(1, min(e2e_steps, 3)), world, dev)
    e2e_val = total_frames / e2e_s

    # ---- BASELINE configs[3] in the same invocation: the fleet, strong scaling
    fleet_block = None
    if args.workload != "fleet" and not args.no_fleet:
        del runner
        torch.cuda.empty_cache()
        fwl = build_rank_workload("fleet", rank, world, 0, args.features, weighted=True)
        frun = make_runner(eng, fwl, rank, dev)
        fl = timed_leg(frun, args.steps, args.warmup, world, dev)
        fe2e_s = timed_host_leg(frun.step_host, max(1, min(args.steps, 5)), world, dev)
        fhist = status_hist(frun)
        Mf = int(fwl["batch"].offsets[-1])
        fleet_block = {"workload": WORKLOADS["fleet"][2], "scaling": "strong", "frames_total": fwl["total_frames"],
                       "frames_this_rank": fwl["n_frames"], "correspondences_this_rank": Mf,
                       "shards": "contiguous frame ranges balanced by correspondences",
                       "value": fwl["total_frames"] / (fl["ms_step"] * 1e-3), "unit": "frames/s", "ms_per_step": fl["ms_step"],
                       "kernel_ms": fl["kernel_ms"], "gather_ms": fl["gather_ms"], "filter_ms": fl["filter_ms"],
                       "gpu_launches": int(fl["launches"]),
                       "e2e": {"value": fwl["total_frames"] / fe2e_s, "unit": "frames/s", "h2d_bytes_per_step": frun.h2d_bytes,
                               "d2h_bytes_per_step": frun.d2h_bytes, "api": "fleet.ShardRunner.step_host (pinned host buffers, copies inside)"},
                       "status_hist": fhist}
        del frun

    # ---- the same kernel on other feature densities (rank 0's GPU, one short sequence each): the bench line's own workload is
    #      image-uniform, which is the easy case for any spatial index
    density_block = None
    if rank == 0 and args.workload == "kitti00" and not args.no_densities:
        density_block = {}
        for name in ("kitti00-ground", "kitti00-clustered"):
            dwl = build_rank_workload(name, 0, 1, 1184, args.features)
            drun = make_runner(eng, dwl, 0, dev)
            dl = timed_leg(drun, max(3, min(args.steps, 5)), 3, 1, dev)
            density_block[name] = {"workload": WORKLOADS[name][2], "frames": 1184, "value": 1184 / (dl["ms_step"] * 1e-3), "unit": "frames/s (1 GPU)",
                                   "kernel_ms": dl["kernel_ms"], "status_hist": status_hist(drun)}
            del drun
    if world > 1:
        dist.barrier()

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.isfile(peaks_path):
            peak = float(json.load(open(peaks_path))["hbm_gbs"]); peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak = 6650.0; peak_src = "fallback (B200_PROFILING.md 6.65 TB/s)"
        M = int(batch.offsets[-1])
        alg_bytes = 16.0 * M + 64.0 * n_frames
        k_ms = main["kernel_ms"]
        achieved = alg_bytes / (k_ms * 1e-3) / 1e9
        # counters of the dominant kernel from the committed ncu capture -- only if it was taken from THESE kernel sources
        traffic = winstr = issue_frac = None
        counters_note = "no ncu capture of this workload committed"
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        src_hash = kernel_source_hash()
        if os.path.isfile(tp) and args.workload == "kitti00" and n_frames == N_FRAMES and n_corr == N_CORR:
            try:
                tj = json.load(open(tp))
                if tj.get("kernel_source_hash") == src_hash:
                    traffic = float(tj["frame_kernel_dram_bytes_per_launch"])
                    winstr = float(tj["frame_kernel_warp_instructions_per_launch"]) / n_frames
                    sm_hz = 1e6 * float((main["clocks"] or {}).get("sm_mhz") or 1965.0)
                    issue_frac = winstr * n_frames / (k_ms * 1e-3) / (148 * 4 * sm_hz)
                    counters_note = "ncu capture %s (scripts/measure_traffic.sh), same kernel sources" % tj.get("captured", "?")
                else:
                    counters_note = "profiles/traffic.json was captured from other kernel sources (hash %s, now %s): refused" % (tj.get("kernel_source_hash"), src_hash)
            except Exception as ex:
                counters_note = "profiles/traffic.json unreadable: %r" % (ex,)
        cpu = cpu_baseline_single(args.cpu_sample, args.workload, args.features) if args.cpu_sample > 0 else None
        dropin = None
        if args.dropin_frames > 0 and args.workload != "dense":
            try:
                dropin = dropin_leg(eng, batch, args.dropin_frames, dev)
            except Exception as ex:
                dropin = {"value": None, "error": repr(ex)}
        mv = batch.move_flags.astype(bool)
        lo = wl["shards"][rank][0]
        mine = scales[lo: lo + n_frames]
        err = np.abs(mine[mv] - batch.true_scale[mv]) / batch.true_scale[mv]
        line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": main["ms_step"], "higher_is_better": True, "scaling": "strong" if args.workload == "fleet" else "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": wl["desc"], "frames_total": total_frames,
                           "frames_per_gpu": n_frames, "correspondences_per_frame": n_corr, "camera": "1241x376", "camera_height_m": 1.7,
                           "outlier_frac": 0.10, "ransac_iterations": 100,
                           "l2": "inputs %.0f MB per pass > 126 MB L2" % (16.0 * M / 1e6),
                           "parallelism": "frame-range shards, %d GPU(s), one in-place all-gather of 16-byte frame records" % world,
                           "status_hist": hist,
                           "scale_rel_err_vs_truth_median": float(np.median(err)), "scale_rel_err_vs_truth_p95": float(np.percentile(err, 95))},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                             "kernel": "frame_kernel<FROM_CORR>", "kernel_ms": k_ms, "algorithmic_bytes_per_launch": alg_bytes,
                             "peak_source": peak_src,
                             "issue_frac": issue_frac, "winstr_per_frame": winstr, "kernel_source_hash": src_hash, "counters": counters_note,
                             "note": "instruction-issue-bound path by construction (SURVEY 8d: 1e5 fps is 0.06 % of the HBM ceiling): issue_frac = "
                                     "warp-instructions / kernel time / (148 SMs x 4 schedulers x SM clock); see profiles/README.md"},
                "phases_ms": {"kernel": k_ms, "gather": main["gather_ms"], "filter": main["filter_ms"]},
                "cpu_baseline": cpu,
                "e2e": {"value": e2e_val, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "api": e2e_api,
                        "pageable_value": (total_frames / page_s) if page_s else None, "dropin_fps": dropin},
                "fleet": fleet_block,
                "other_densities": density_block,
                "gpu_launches": int(main["launches"]), "clocks": main["clocks"]}
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="kitti00", choices=sorted(WORKLOADS), help="kitti00 = BASELINE configs[1] (the bench line); dense / fleet = configs[2] / configs[3]; kitti00-ground / kitti00-clustered: perspective and clustered feature densities")
    ap.add_argument("--frames", type=int, default=0, help="frames per GPU (0 = the workload's own)")
    ap.add_argument("--features", type=int, default=0, help="correspondences per frame (0 = the workload's own)")
    ap.add_argument("--cpu-sample", type=int, default=50, help="frames of the workload timed on one host core through the reference (0 = skip)")
    ap.add_argument("--dropin-frames", type=int, default=500, help="frames of the per-frame drop-in leg (0 = skip)")
    ap.add_argument("--no-densities", action="store_true", help="skip the perspective / clustered feature-density legs of the kitti00 line")
    ap.add_argument("--no-fleet", action="store_true", help="skip the fleet block (BASELINE configs[3] in the same invocation)")
    ap.add_argument("--ref-cores", type=int, default=0, help="--impl reference: worker processes (0 = all host cores, at most 64)")
    ap.add_argument("--ref-frames", type=int, default=0, help="--impl reference: frames per worker per step (0 = 2 for the reference, 6 for the port)")
    ap.add_argument("--ref-port", action="store_true", help="--impl reference: time the oracle port even when the reference is staged")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
