"""numpy-in / numpy-out helpers over libmvosr.so for the drop-in modules of this directory.  Every function launches CUDA
kernels through the C ABI (include/mvosr.h); without a CUDA device they raise -- there is no CPU fallback."""
import numpy as np

_ENGINES = {}


def engine(absolute_reference=1.75, vanish=185.0):
    """One native handle per (camera height, ROI row) pair, created on first use."""
    from mvoscalerecovery_b200.batch import ScaleRecovery
    key = (float(absolute_reference), float(vanish))
    if key not in _ENGINES:
        _ENGINES[key] = ScaleRecovery(absolute_reference=key[0], vanish=key[1])
    return _ENGINES[key]


def _dev(eng, a, dtype):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a, dtype=dtype)).to(eng.device)


def delaunay(points2d):
    """Canonical Delaunay simplices (rows ascending, rows lexsorted) of an (n,2) point set: the replacement of
    scipy.spatial.Delaunay(points).simplices (src/rescale.py:124-125, src/scale_calculator.py:257,266).  Coordinates are
    taken as float32 (the kernel's exact predicates are defined on float32 pixels)."""
    eng = engine()
    p = np.asarray(points2d, dtype=np.float64).reshape(-1, 2)
    n = p.shape[0]
    off = np.array([0, n], np.int32)
    out = eng.delaunay_frames(_dev(eng, off, np.int32), _dev(eng, p[:, 0], np.float32), _dev(eng, p[:, 1], np.float32), max(n, 1))
    nt = int(out["n_tri"].cpu().numpy()[0])
    st = int(out["status"].cpu().numpy()[0])
    if st & ~2:
        raise RuntimeError("Delaunay failed: per-frame status 0x%02x (fewer than 3 points, all collinear, or out of range)" % st)
    return out["tri"][:nt].cpu().numpy()


def triangle_planes(triangle_ids, feature3d):
    """(normal (T,3) = P^-1 1, height (T,) = 1/|n|, mean_y (T,)) of caller-supplied triangles (src/rescale.py:77-84)."""
    eng = engine()
    tri = np.asarray(triangle_ids).reshape(-1, 3)
    if tri.shape[0] == 0:
        return np.zeros((0, 3)), np.zeros(0), np.zeros(0)
    out = eng.triangle_planes(_dev(eng, tri, np.int32), _dev(eng, np.asarray(feature3d).reshape(-1, 3), np.float64))
    return out["normal"].cpu().numpy(), out["height"].cpu().numpy(), out["mean_y"].cpu().numpy()


def triangle_votes(triangle_ids, pixel_v, depth):
    """(flagged, incident) per vertex as float64 arrays (check_triangle + find_outliers, src/rescale.py:45-72)."""
    eng = engine()
    tri = np.asarray(triangle_ids).reshape(-1, 3)
    out = eng.triangle_votes(_dev(eng, tri, np.int32), _dev(eng, pixel_v, np.float64), _dev(eng, depth, np.float64))
    return out["flagged"].cpu().numpy().astype(np.float64), out["incident"].cpu().numpy().astype(np.float64)


def ransac_plane(points, max_iterations, threshold, seed=0, frame=0, seq=0, goal_fraction=0.8, stop_at_goal=True):
    """get_pitch_ransac on the GPU (src/estimate_road_norm.py:66-70): (model (4,), inlier count, best hypothesis, used)."""
    eng = engine()
    pts = np.asarray(points, dtype=np.float64).reshape(-1, 3)
    off = np.array([0, pts.shape[0]], np.int32)
    out = eng.ransac_planes(_dev(eng, off, np.int32), _dev(eng, pts, np.float64), iterations=int(max_iterations), threshold=float(threshold),
                            goal_fraction=goal_fraction, stop_at_goal=stop_at_goal, seed=seed,
                            frame_index=_dev(eng, np.array([frame], np.int32), np.int32), seq_id=seq)
    m = out["model"].cpu().numpy()[0]
    return m, int(out["ic"].cpu().numpy()[0]), int(out["best_hyp"].cpu().numpy()[0]), int(out["hyps_used"].cpu().numpy()[0])


def integrate_path(motions, scales=None):
    """get_path / motion2pose (src/main_offline.py:95-119): (F,12) motions + (F,) scales -> (F+1,12) poses."""
    eng = engine()
    mot = np.asarray(motions, dtype=np.float64).reshape(-1, 12)
    off = np.array([0, mot.shape[0]], np.int32)
    sc = None if scales is None else _dev(eng, np.asarray(scales, dtype=np.float64).reshape(-1), np.float64)
    return eng.integrate_paths(_dev(eng, off, np.int32), _dev(eng, mot, np.float64), sc).cpu().numpy()


def depth_from_mesh(cam, triangle_ids, points2d, datas):
    """Reconstruct.depth_generate (src/reconstruct.py:91-107): (depth (H,W), tri_id (H,W)) for a camera with
    width/height/fx/fy/cx/cy attributes."""
    eng = engine()
    out = eng.depth_from_mesh(int(cam.width), int(cam.height), cam.fx, cam.fy, cam.cx, cam.cy,
                              _dev(eng, np.asarray(triangle_ids).reshape(-1, 3), np.int32),
                              _dev(eng, np.asarray(points2d).reshape(-1, 2), np.float64), _dev(eng, np.asarray(datas).reshape(-1, 4), np.float64))
    return out["depth"].cpu().numpy(), out["tri_id"].cpu().numpy()
