"""CPU restatements of the steps around the hot path -- TEST INFRASTRUCTURE (oracle side).

Part of ``oracle/``: imported only by ``tests/`` (and allowed for ``__graft_entry__.smoke()`` / ``bench.py``'s CPU legs).  The
product has its own device implementations (csrc/aux_kernels.cuh); these functions restate what the reference (or the
third-party library it calls) computes, in plain numpy, each citing the lines it follows, and are pinned by
``tests/test_oracle_extras.py`` against outputs of the reference / of OpenCV generated in the build container
(tests/golden/scripts.npz, pose.npz).
"""
from __future__ import annotations

import numpy as np

from . import pipeline as P


def motion2pose(motions, scales=None):
    """get_path + motion2pose of src/main_offline.py:95-119: translations scaled by the per-frame scale, then the running
    left-to-right product of the 4x4 relative motions; row 0 is the identity.  (N,12) -> (N+1,12)."""
    mot = np.array(motions, dtype=np.float64).reshape(-1, 12)
    if scales is not None:
        mot[:, 3:12:4] = (np.asarray(scales, dtype=np.float64) * mot[:, 3:12:4].T).T
    poses = np.zeros((mot.shape[0] + 1, 12))
    cur = np.eye(4)
    poses[0] = cur[:3].reshape(-1)
    for i in range(mot.shape[0]):
        m = np.eye(4)
        m[:3] = mot[i].reshape(3, 4)
        cur = cur @ m
        poses[i + 1] = cur[:3].reshape(-1)
    return poses


def decompose_essential(E):
    """cv::decomposeEssentialMat (OpenCV 4.x calib3d/five-point.cpp), the call inside cv2.recoverPose that
    src/thirdparty/MonocularVO/visual_odometry.py:132 makes: SVD, det(U) = det(Vt) = +1, R1 = U W Vt, R2 = U W^T Vt, t = U[:,2]."""
    U, _, Vt = np.linalg.svd(np.asarray(E, dtype=np.float64).reshape(3, 3))
    if np.linalg.det(U) < 0:
        U = -U
    if np.linalg.det(Vt) < 0:
        Vt = -Vt
    W = np.array([[0.0, 1.0, 0.0], [-1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])
    return U @ W @ Vt, U @ W.T @ Vt, U[:, 2].copy()


def recover_pose(E, cur_uv, ref_uv, fx, fy, cx, cy, dist=100.0):
    """cv2.recoverPose(E, px_cur, px_ref, K, distanceThresh=dist) (visual_odometry.py:129-133): the four (R, t) candidates of
    the decomposition, every correspondence triangulated under each (the DLT of oracle.pipeline.triangulate_dlt with
    recoverPose's mask), the candidate with the most points in front of both cameras and nearer than ``dist`` wins, the
    first one on ties in OpenCV's order (R1,t), (R2,t), (R1,-t), (R2,-t).  Returns (R, t, mask, counts)."""
    R1, R2, t = decompose_essential(E)
    cands = [(R1, t), (R2, t), (R1, -t), (R2, -t)]
    masks = [P.triangulate_dlt(cur_uv, ref_uv, R, tt, fx, fy, cx, cy, dist)[1] for R, tt in cands]
    counts = [int(m.sum()) for m in masks]
    best = 0
    for k in range(1, 4):                       # good1 >= good2 && ... : the first maximum
        if counts[k] > counts[best]:
            best = k
    return cands[best][0], cands[best][1], masks[best], counts


def triangle_planes(feature3d, triangle_ids):
    """The loop body of flat_selection (src/rescale.py:77-84) / feature_selection_by_tri (src/scale_calculator.py:228-237):
    n = P^-1 1 by LAPACK as the reference does (np.matrix(...).I), height = 1/|n|, mean Y of the vertices."""
    f3 = np.asarray(feature3d, dtype=np.float64)
    tri = np.asarray(triangle_ids).reshape(-1, 3)
    n = np.stack([np.linalg.inv(f3[t]) @ np.ones(3) for t in tri]) if tri.shape[0] else np.zeros((0, 3))
    return n, 1.0 / np.sqrt(np.sum(n * n, 1)), f3[tri][:, :, 1].mean(1) if tri.shape[0] else np.zeros(0)


def triangle_votes(triangle_ids, pixel_v, depth, n_points):
    """check_triangle + find_outliers of src/rescale.py:45-72: per vertex the number of triangles flagging it ([a|b, a|b|c, c]
    with a, b, c the (v_i-v_j)(d_i-d_j) > 0 tests of edges 01, 02, 12) and the number of incident triangles."""
    tri = np.asarray(triangle_ids).reshape(-1, 3)
    v, d = np.asarray(pixel_v)[tri], np.asarray(depth)[tri]
    a = (v[:, 0] - v[:, 1]) * (d[:, 0] - d[:, 1]) > 0
    b = (v[:, 0] - v[:, 2]) * (d[:, 0] - d[:, 2]) > 0
    c = (v[:, 1] - v[:, 2]) * (d[:, 1] - d[:, 2]) > 0
    flags = np.stack([a | b, a | b | c, c], 1)
    return np.bincount(tri[flags], minlength=n_points), np.bincount(tri.reshape(-1), minlength=n_points)


def depth_from_mesh(width, height, fx, fy, cx, cy, delaunay, datas):
    """Reconstruct.depth_generate (src/reconstruct.py:91-107): tri.find_simplex of every integer pixel and the depth of that
    triangle's plane along the pixel's ray, h / (n . ((u-cx)/fx, (v-cy)/fy, 1)); 0 outside.  Returns (depth (H,W), ids (H,W))."""
    v, u = np.mgrid[0:height, 0:width]
    ids = delaunay.find_simplex(np.stack([u.ravel(), v.ravel()], 1).astype(np.float64))
    d = np.asarray(datas, dtype=np.float64)[np.maximum(ids, 0)]
    depth = d[:, 3] / (d[:, 0] * ((u.ravel() - cx) / fx) + d[:, 1] * ((v.ravel() - cy) / fy) + d[:, 2])
    depth[ids < 0] = 0.0
    return depth.reshape(height, width), ids.reshape(height, width)


def bucket_philox(features, bucket_size=30, density=2, seed=0, frame=0, seq=0):
    """bucket(features, bucket_size, density) of src/detector.py:65-95 with the shuffle replaced by a DEFINED order: positions of
    the surviving features, cell by cell (rows of cells top to bottom, cells left to right), inside a cell the `density`
    smallest (r_i, i), r_i = word 0 of Philox4x32-10(counter = (i, frame, seq, 3), key = seed).  The reference shuffles with
    numpy's global RNG: which members of a cell survive cannot be reproduced, the cells, their order and the counts can."""
    from .philox import MASK, philox4x32_10
    f = np.asarray(features, dtype=np.float32).reshape(-1, 2)
    key = (seed & MASK, (seed >> 32) & MASK)
    cells = {}
    for i, (u, v) in enumerate(f):
        cells.setdefault((int(v) // bucket_size, int(u) // bucket_size), []).append((philox4x32_10((i, frame, seq, 3), key)[0], i))
    out = []
    for c in sorted(cells):
        out += [i for _, i in sorted(cells[c])[:density]]
    return np.array(out, dtype=np.int32)
