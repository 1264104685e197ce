"""Writes tests/golden/main_c1.npz -- BASELINE configs[0] in small: the reference's own ``src/main.py`` (AKAZE + LK tracking
+ cv2.findEssentialMat / recoverPose + ScaleEstimator), UNMODIFIED, run on a rendered textured-ground-plane sequence
(1241x376, KITTI-00 intrinsics, camera 1.75 m above the plane = param.camera_h, known forward steps), and what the reference
then makes of its own hand-off file.

Run in the build container only (needs /root/reference, OpenCV):  python tests/golden/make_main_golden.py

Steps: (1) render N frames and write the image list / calibration file main.py expects; (2) execute src/main.py from its file
under the harness shims (matplotlib stub, np.float) -- it writes ``result/<name>_result.npy<tag>.npy``, the pickled hand-off
of src/main.py:149-154; (3) read the hand-off the way src/main_offline.py does, round it to float32 (the container's and the
kernels' input type -- the parity protocol of SURVEY.md section 8c: both sides consume identical bit patterns) and run the
reference's estimator over it through oracle/ref_harness.py (Philox hypothesis stream, canonicalised Delaunay): the golden
scales.  The scales the reference produced on the unrounded float64 features are stored too, for information.
"""
import contextlib
import io
import os
import runpy
import sys
import tempfile

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness as H                      # noqa: E402
from mvoscalerecovery_b200 import container              # noqa: E402

SEED = 31337
N_FRAMES = 40
W, HGT, F, CX, CY = 1241, 376, 718.856, 607.1928, 185.2157
CAM_H = 1.75


def make_texture(rng, size=4096):
    """A corner-rich ground texture: random rectangles and discs over multi-scale noise."""
    tex = np.zeros((size, size), np.float32)
    for sigma, amp in ((64, 60.0), (16, 50.0), (4, 40.0)):
        n = rng.standard_normal((size, size)).astype(np.float32)
        tex += amp * cv2.GaussianBlur(n, (0, 0), sigma) * sigma
    tex = (tex - tex.min()) / (tex.max() - tex.min()) * 160 + 40
    img = tex.astype(np.uint8)
    for _ in range(9000):
        x, y = rng.integers(0, size, 2)
        s = int(rng.integers(4, 40))
        c = int(rng.integers(0, 256))
        if rng.random() < 0.5:
            cv2.rectangle(img, (int(x), int(y)), (int(x + s), int(y + int(s * rng.uniform(0.4, 1.6)))), c, -1)
        else:
            cv2.circle(img, (int(x), int(y)), s // 2, c, -1)
    return cv2.GaussianBlur(img, (0, 0), 1.0)


def render(tex, cam_x, cam_z, yaw, metres_per_texel=0.02):
    """Image of the plane Y = CAM_H (camera frame, Y down) seen from (cam_x, cam_z) with heading yaw; sky rows flat grey."""
    v, u = np.mgrid[0:HGT, 0:W].astype(np.float32)
    below = v > CY + 2
    Z = np.where(below, F * CAM_H / np.maximum(v - CY, 1e-3), 0)
    X = (u - CX) * Z / F
    xw = cam_x + np.cos(yaw) * X + np.sin(yaw) * Z
    zw = cam_z - np.sin(yaw) * X + np.cos(yaw) * Z
    size = tex.shape[0]
    mx = np.mod(xw / metres_per_texel + size / 2, size - 1).astype(np.float32)
    my = np.mod(zw / metres_per_texel, size - 1).astype(np.float32)
    img = cv2.remap(tex, mx, my, cv2.INTER_LINEAR)
    img[~below] = 128
    return cv2.cvtColor(img, cv2.COLOR_GRAY2BGR)


def main():
    rng = np.random.default_rng(SEED)
    tex = make_texture(rng)
    steps = 0.8 + 0.15 * np.sin(np.arange(N_FRAMES) / 5.0)
    ns = H.load_reference(seed=SEED)
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "dataset")); os.makedirs(os.path.join(tmp, "result")); os.makedirs(os.path.join(tmp, "img"))
        open(os.path.join(tmp, "dataset", "00_calib.txt"), "w").write("P0: %r 0 %r 0 0 %r %r 0 0 0 1 0\n" % (F, CX, F, CY))
        names, x, z, yaw = [], 0.0, 0.0, 0.0
        for i in range(N_FRAMES):
            name = os.path.join(tmp, "img", "%06d.png" % i)
            cv2.imwrite(name, render(tex, x, z, yaw))
            names.append(name)
            yaw += 0.002 * np.sin(i / 7.0)
            x += steps[i] * np.sin(yaw); z += steps[i] * np.cos(yaw)
        lst = os.path.join(tmp, "synth_00.txt")
        open(lst, "w").write("header line (main.py skips the first line)\n" + "\n".join(names) + "\n")
        old_argv, old_cwd = sys.argv, os.getcwd()
        try:
            sys.argv = ["main.py", lst, ".g"]
            os.chdir(tmp)
            with contextlib.redirect_stdout(io.StringIO()):
                runpy.run_path(os.path.join(H.REF_SRC, "main.py"), run_name="__main__")
        finally:
            sys.argv = old_argv
            os.chdir(old_cwd)
        handoff = os.path.join(tmp, "result", "synth_00_result.npy.g.npy")
        ref_scales_f64 = np.loadtxt(os.path.join(tmp, "result", "synth_00_scales.txt.g"))
        seq = container.load_reference_npy(handoff)                     # float32 structure-of-arrays, as the batch path consumes it
        raw = np.load(handoff, allow_pickle=True).item()
    lists = container.unpack_sequence(seq)                               # float64 views of the float32 values
    scales, recs = H.run_offline_loop(ns, lists["feature3ds"], lists["feature2ds"], lists["move_flags"],
                                      absolute_reference=ns.param.camera_h, window_size=5, seq=0)
    n_feat = np.diff(seq["offsets"])
    out = dict(seed=np.uint64(SEED), absolute_reference=np.float64(ns.param.camera_h), true_steps=steps, scales=scales,
               ref_main_scales_unrounded=ref_scales_f64, called=np.array([r is not None for r in recs]),
               raw_scale=np.array([r["raw_scale"] if r is not None else np.nan for r in recs]),
               n_kept=np.array([int(r["keep"].sum()) if r is not None else -1 for r in recs]),
               n_tri=np.array([r["tri2"].shape[0] if r is not None else -1 for r in recs]),
               best_ic=np.array([r["best_ic"] if r is not None else -1 for r in recs]), **{k: np.asarray(v) for k, v in seq.items()})
    path = os.path.join(ROOT, "tests", "golden", "main_c1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.0f KB" % (os.path.getsize(path) / 1024), "frames", len(recs), "features/frame", n_feat.min(), n_feat.max())
    print("moving", int(seq["move_flags"].sum()), "estimator calls", int(out["called"].sum()))
    print("scales (harness, f32 hand-off)", np.round(scales[:12], 4))
    print("scales (main.py itself, f64, OS-entropy RANSAC)", np.round(ref_scales_f64[:12], 4))
    print("true steps", np.round(steps[:12], 4), "len(raw lists)", len(raw["feature3ds"]))


if __name__ == "__main__":
    main()
