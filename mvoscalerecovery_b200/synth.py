"""Synthetic KITTI-shaped sequences: a ground plane seen from a
forward-moving camera 1.7 m above it, 1241x376 images, known relative poses.

One *frame* is the set of tracked correspondences between the previous ("ref")
and the current image plus the relative pose (R, t) with ``x_ref = R x_cur + t``
and ``|t| = 1`` (the convention of ``cv2.recoverPose`` as used by the reference,
src/thirdparty/MonocularVO/visual_odometry.py:129-147).  Everything is in VO
units (metres divided by the frame's true step length s_f), so the scale the
pipeline should recover for frame f is s_f.

The generator is numpy-only, deterministic in (seed, seq, frame) and shared by
tests, goldens, bench.py and the CPU baseline so both sides see identical bits.
Pixel arrays are float32: that is the bit pattern both the CUDA path and the
reference are fed ("inputs must be identical bit patterns").
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

# KITTI odometry sequence lengths 00..10
KITTI_SEQ_LENGTHS = (4541, 1101, 4661, 801, 271, 2761, 1101, 1101, 4071, 1591, 1201)


@dataclass(frozen=True)
class Camera:
    """Pinhole camera of src/param.py:30-35 (KITTI 00)."""
    width: float = 1241.0
    height: float = 376.0
    fx: float = 718.856
    fy: float = 718.856
    cx: float = 607.1928
    cy: float = 185.2157


@dataclass
class CorrespondenceBatch:
    """CSR-packed correspondences for F frames (structure-of-arrays, float32)."""
    offsets: np.ndarray      # (F+1,) int32
    cur_u: np.ndarray        # (M,) f32 pixel in current image
    cur_v: np.ndarray
    ref_u: np.ndarray        # (M,) f32 pixel in previous image
    ref_v: np.ndarray
    poses: np.ndarray        # (F,12) f64, row-major [R|t] ("motion" rows of main.py:129-133)
    move_flags: np.ndarray   # (F,) uint8
    true_scale: np.ndarray   # (F,) f64 step length in metres

    @property
    def n_frames(self) -> int:
        return int(self.offsets.shape[0] - 1)


def true_scale_profile(n_frames: int, seed: int, seq: int = 0) -> np.ndarray:
    """Smooth KITTI-like step lengths: median ~0.85 m, range 0.25..1.45, |ds| << 0.3."""
    rng = np.random.Generator(np.random.Philox(key=[seed & 0xFFFFFFFFFFFFFFFF, 0x5CA1E + seq]))
    f = np.arange(n_frames, dtype=np.float64)
    ph = rng.uniform(0, 2 * np.pi, size=3)
    s = (0.85 + 0.38 * np.sin(2 * np.pi * f / 523.0 + ph[0])
         + 0.16 * np.sin(2 * np.pi * f / 131.0 + ph[1])
         + 0.05 * np.sin(2 * np.pi * f / 23.0 + ph[2]))
    return np.clip(s, 0.25, 1.45)


def _rodrigues(rx, ry, rz):
    th = np.sqrt(rx * rx + ry * ry + rz * rz)
    if th < 1e-12:
        return np.eye(3)
    k = np.array([rx, ry, rz]) / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)


def make_frame(seed: int, seq: int, frame: int, n_corr: int, true_scale: float,
               cam: Camera = Camera(), camera_h: float = 1.7, outlier_frac: float = 0.1,
               upper_frac: float = 0.2, pixel_noise: float = 0.05, z_max: float = 60.0, density: str = "uniform"):
    """One frame of correspondences.

    Returns (cur_uv (n,2) f32, ref_uv (n,2) f32, R (3,3) f64, t (3,) f64).

    Scene: road points are sampled uniformly over the image region below the
    horizon (real front-ends bucket features, src/detector.py:65-95, so image-uniform
    is the realistic density), back-projected onto the plane n.X = h (h = camera_h /
    true_scale in VO units, n tilted by <1 deg of pitch/roll).  ``outlier_frac`` of
    them are lifted 0.2..1.5 m off the road (kerbs, cars) or given a corrupted
    depth; ``upper_frac`` of all correspondences lie above the horizon (facades)
    and are removed by the ROI cut (src/rescale.py:115).

    ``density``: where the road features fall.  "uniform" -- uniform in the image (bucketed front-end, the default);
    "ground" -- uniform on the GROUND, X in U(-8, 8) m, Z in U(5, 40) m (SURVEY.md section 8d verbatim): in the image the
    density grows like 1/(v - cy)^3 towards the horizon; "clustered" -- 70 % of them in 12 Gaussian clusters (textured
    patches: sigma 30 x 9 px) over a uniform background.
    """
    rng = np.random.Generator(np.random.Philox(key=[seed & 0xFFFFFFFFFFFFFFFF,
                                                    (seq << 32) | (frame & 0xFFFFFFFF)]))
    h = camera_h / true_scale
    n_up = int(round(n_corr * upper_frac))
    n_low = n_corr - n_up
    # plane normal in the camera frame (Y down), small pitch / roll
    pitch = np.deg2rad(rng.uniform(-0.8, 0.8))
    roll = np.deg2rad(rng.uniform(-0.5, 0.5))
    nrm = _rodrigues(pitch, 0.0, roll) @ np.array([0.0, 1.0, 0.0])

    # ---- lower (road ROI) points: uniform in the image below the far-depth row
    v_top = cam.cy + cam.fy * h / z_max
    v_top = min(max(v_top, cam.cy + 6.0), cam.height - 40.0)
    if density == "uniform":
        u = rng.uniform(0.0, cam.width - 1.0, size=n_low)
        v = rng.uniform(v_top, cam.height - 1.0, size=n_low)
    elif density == "ground":
        u = np.empty(0); v = np.empty(0)
        while u.shape[0] < n_low:                               # rejection: keep what projects inside the image below the far row
            Xg = rng.uniform(-8.0, 8.0, size=2 * n_low) / true_scale
            Zg = rng.uniform(5.0, 40.0, size=2 * n_low) / true_scale
            Yg = (h - nrm[0] * Xg - nrm[2] * Zg) / nrm[1]
            ug, vg = Xg / Zg * cam.fx + cam.cx, Yg / Zg * cam.fy + cam.cy
            ok = (ug >= 0.0) & (ug <= cam.width - 1.0) & (vg >= v_top) & (vg <= cam.height - 1.0)
            u, v = np.concatenate([u, ug[ok]]), np.concatenate([v, vg[ok]])
        u, v = u[:n_low], v[:n_low]
    elif density == "clustered":
        n_cl = int(round(0.7 * n_low))
        cu = rng.uniform(40.0, cam.width - 41.0, size=12)
        cv = rng.uniform(v_top + 10.0, cam.height - 11.0, size=12)
        which = rng.integers(0, 12, size=n_cl)
        u = np.concatenate([cu[which] + 30.0 * rng.standard_normal(n_cl), rng.uniform(0.0, cam.width - 1.0, size=n_low - n_cl)])
        v = np.concatenate([cv[which] + 9.0 * rng.standard_normal(n_cl), rng.uniform(v_top, cam.height - 1.0, size=n_low - n_cl)])
        u, v = np.clip(u, 0.0, cam.width - 1.0), np.clip(v, v_top, cam.height - 1.0)
    else:
        raise ValueError("density must be 'uniform', 'ground' or 'clustered'")
    rx = (u - cam.cx) / cam.fx
    ry = (v - cam.cy) / cam.fy
    denom = nrm[0] * rx + nrm[1] * ry + nrm[2]
    denom = np.where(denom > 1e-3, denom, 1e-3)
    z = h / denom
    kind = rng.uniform(size=n_low)
    lifted = kind < outlier_frac * 0.6
    corrupt = (kind >= outlier_frac * 0.6) & (kind < outlier_frac)
    lift_m = rng.uniform(0.2, 1.5, size=n_low) / true_scale           # VO units
    z_l = (h - lift_m) / denom
    z = np.where(lifted & (z_l > 0.5), z_l, z)
    z = np.where(corrupt, z * rng.uniform(0.5, 2.0, size=n_low), z)
    z = np.clip(z, 0.5, 95.0)
    low = np.stack([rx * z, ry * z, z], axis=1)

    # ---- upper points (facades / sky-line): removed by the ROI cut
    uu = rng.uniform(0.0, cam.width - 1.0, size=n_up)
    vu = rng.uniform(0.0, cam.cy - 2.0, size=n_up)
    zu = rng.uniform(5.0, 80.0, size=n_up)
    up = np.stack([(uu - cam.cx) / cam.fx * zu, (vu - cam.cy) / cam.fy * zu, zu], axis=1)

    X = np.concatenate([low, up], axis=0)
    perm = rng.permutation(X.shape[0])          # interleave road / facade features
    X = X[perm]

    # ---- relative pose: mostly forward, small yaw / pitch, unit translation
    R = _rodrigues(np.deg2rad(rng.uniform(-0.3, 0.3)), np.deg2rad(rng.uniform(-1.0, 1.0)),
                   np.deg2rad(rng.uniform(-0.2, 0.2)))
    t = np.array([rng.uniform(-0.05, 0.05), rng.uniform(-0.02, 0.02), 1.0])
    t = t / np.linalg.norm(t)

    Xr = X @ R.T + t
    cur = np.stack([X[:, 0] / X[:, 2] * cam.fx + cam.cx, X[:, 1] / X[:, 2] * cam.fy + cam.cy], axis=1)
    ref = np.stack([Xr[:, 0] / Xr[:, 2] * cam.fx + cam.cx, Xr[:, 1] / Xr[:, 2] * cam.fy + cam.cy], axis=1)
    ref = ref + rng.normal(0.0, pixel_noise, size=ref.shape)
    return cur.astype(np.float32), ref.astype(np.float32), R, t


def sequence_sizes(seed: int, n_frames: int, n_corr: int = 2500, seq: int = 0, n_jitter: float = 0.05) -> np.ndarray:
    """Correspondences per frame of ``make_sequence(seed, n_frames, n_corr, seq, n_jitter=...)`` without building the frames
    (still frames aside): what a fleet scheduler needs to balance shards by work before any rank generates its data."""
    rng = np.random.Generator(np.random.Philox(key=[seed & 0xFFFFFFFFFFFFFFFF, 0xC0FFEE + seq]))
    return np.maximum(8, np.round(n_corr * (1 + rng.uniform(-n_jitter, n_jitter, n_frames)))).astype(np.int64)


def make_sequence(seed: int, n_frames: int, n_corr: int = 2500, seq: int = 0,
                  cam: Camera = Camera(), camera_h: float = 1.7, outlier_frac: float = 0.1,
                  n_jitter: float = 0.05, still_every: int = 0, scales=None, frame_range=None, **kw) -> CorrespondenceBatch:
    """A CSR batch of ``n_frames`` frames; frame sizes jitter by +-n_jitter around n_corr.
    ``frame_range=(f0, f1)`` builds only frames f0..f1-1 of that sequence (same content as in the full sequence).

    ``still_every`` > 0 marks every k-th frame as "not moving" (move_flag 0, no
    correspondences), the case of src/main_offline.py:64-68.
    """
    scales = true_scale_profile(n_frames, seed, seq) if scales is None else np.asarray(scales, dtype=np.float64)
    sizes = sequence_sizes(seed, n_frames, n_corr, seq, n_jitter)
    move = np.ones(n_frames, dtype=np.uint8)
    if still_every:
        move[still_every - 1::still_every] = 0
        sizes[move == 0] = 0
    f0, f1 = (0, n_frames) if frame_range is None else (int(frame_range[0]), int(frame_range[1]))
    sizes, move, scales, first = sizes[f0:f1], move[f0:f1], scales[f0:f1], f0
    n_frames = f1 - f0
    offsets = np.zeros(n_frames + 1, dtype=np.int64)
    np.cumsum(sizes, out=offsets[1:])
    M = int(offsets[-1])
    cu = np.empty(M, np.float32); cv_ = np.empty(M, np.float32)
    ru = np.empty(M, np.float32); rv = np.empty(M, np.float32)
    poses = np.zeros((n_frames, 12), np.float64)
    poses[:, [0, 5, 10]] = 1.0
    for f in range(n_frames):
        if not move[f]:
            continue
        c, r, R, t = make_frame(seed, seq, first + f, int(sizes[f]), float(scales[f]), cam, camera_h,
                                outlier_frac, **kw)
        a, b = offsets[f], offsets[f + 1]
        cu[a:b] = c[:, 0]; cv_[a:b] = c[:, 1]; ru[a:b] = r[:, 0]; rv[a:b] = r[:, 1]
        P = np.zeros((3, 4)); P[:, :3] = R; P[:, 3] = t
        poses[f] = P.reshape(-1)
    return CorrespondenceBatch(offsets.astype(np.int32), cu, cv_, ru, rv, poses, move, scales)
