"""BASELINE configs[0] timed: the reference's UNMODIFIED src/main.py (AKAZE + LK tracking + findEssentialMat / recoverPose +
ScaleEstimator per frame, src/main.py:55,87-113) on a rendered textured-ground-plane sequence (1241x376, KITTI-00 intrinsics,
camera 1.75 m above the plane, known forward steps), whole-program wall clock,
  (a) with the reference's own modules (rescale.py, graph.py, estimate_road_norm.py ... numpy / scipy / Python loops), and
  (b) with mvoscalerecovery_b200/compat first on the module path: the same main.py, the same front-end,
      rescale.ScaleEstimator.scale_calculation = one C-ABI call into the fused frame kernel per frame.
Each arm is a fresh interpreter (imports, CUDA context creation and the first launch are inside arm (b)'s cold figure); arm (b) then
runs main.py a second time in the same process for the steady-state figure.  The front-end is identical in both arms and runs on the
host: what the drop-in removes is the estimator's share.  The reference arm runs on the first --ref-frames frames (it needs ~0.6 s a
frame; say so in the output).

  python scripts/time_main_config0.py --frames 500 --ref-frames 100        (GPU box: needs cuda:0 and oracle/_ref or /root/reference)

Prints one JSON object; rendered frames and result files live in a temporary directory."""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

CHILD = r'''
import contextlib, io, json, os, runpy, sys, time, types
import numpy as np
ref_src, compat, root, lst, tag, repeats = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5], int(sys.argv[6])
for m in ("matplotlib", "matplotlib.pyplot"):                      # the harness shims of oracle/ref_harness.py: plotting stub, removed numpy alias
    sys.modules.setdefault(m, types.ModuleType(m))
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
if not hasattr(np, "float"):
    np.float = float
sys.path.insert(0, ref_src)
if compat:
    sys.path.insert(0, root); sys.path.insert(0, compat)          # PYTHONPATH=repo:repo/mvoscalerecovery_b200/compat (INTEGRATION.md section 1)
times = []
for r in range(repeats):
    sys.argv = ["main.py", lst, ".r%d" % r + tag]
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        runpy.run_path(os.path.join(ref_src, "main.py"), run_name="__main__")
    times.append(time.perf_counter() - t0)
import rescale
out = {"times": times, "rescale": os.path.abspath(rescale.__file__)}
sc = os.path.join("result", os.path.basename(lst)[:-4] + "_scales.txt.r%d" % (repeats - 1) + tag)
if os.path.isfile(sc):
    out["scales"] = np.loadtxt(sc).tolist()
print("@@" + json.dumps(out))
'''


def run_arm(ref_src, compat, lst, tag, repeats, cwd):
    t0 = time.perf_counter()
    p = subprocess.run([sys.executable, "-c", CHILD, ref_src, compat, ROOT, lst, tag, str(repeats)], cwd=cwd, capture_output=True, text=True)
    wall = time.perf_counter() - t0
    line = [l for l in p.stdout.splitlines() if l.startswith("@@")]
    if p.returncode != 0 or not line:
        raise RuntimeError("main.py arm failed (%s): %s" % (tag, p.stderr[-2000:]))
    out = json.loads(line[0][2:])
    out["process_wall_s"] = wall
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=500)
    ap.add_argument("--ref-frames", type=int, default=100, help="frames of the reference arm (0 = skip it)")
    ap.add_argument("--skip-dropin", action="store_true", help="reference arm only (no GPU needed)")
    args = ap.parse_args()
    import cv2
    import numpy as np
    import make_main_golden as G                               # the renderer of the committed configs[0] golden (tests/golden)
    from oracle import ref_harness as H
    ref_src = H.REF_SRC
    assert os.path.isfile(os.path.join(ref_src, "main.py")), "reference sources not found (make -C oracle ref)"
    compat = os.path.join(ROOT, "mvoscalerecovery_b200", "compat")
    rng = np.random.default_rng(G.SEED)
    tex = G.make_texture(rng)
    steps = 0.8 + 0.15 * np.sin(np.arange(args.frames) / 5.0)
    with tempfile.TemporaryDirectory() as tmp:
        for d in ("dataset", "result", "img"):
            os.makedirs(os.path.join(tmp, d))
        open(os.path.join(tmp, "dataset", "00_calib.txt"), "w").write("P0: %r 0 %r 0 0 %r %r 0 0 0 1 0\n" % (G.F, G.CX, G.F, G.CY))
        names, x, z, yaw = [], 0.0, 0.0, 0.0
        t0 = time.perf_counter()
        for i in range(args.frames):
            name = os.path.join(tmp, "img", "%06d.png" % i)
            cv2.imwrite(name, G.render(tex, x, z, yaw))
            names.append(name)
            yaw += 0.002 * np.sin(i / 7.0)
            x += steps[i] * np.sin(yaw); z += steps[i] * np.cos(yaw)
        render_s = time.perf_counter() - t0

        def listing(n, name):
            path = os.path.join(tmp, name)
            open(path, "w").write("header line (main.py skips the first line)\n" + "\n".join(names[:n]) + "\n")
            return path

        res = {"config": "BASELINE configs[0]: unmodified src/main.py on %d rendered frames 1241x376, camera height %.2f m" % (args.frames, G.CAM_H),
               "render_s": render_s}
        d = None
        if not args.skip_dropin:
            d = run_arm(ref_src, compat, listing(args.frames, "synth_00.txt"), ".dropin", 2, tmp)
        if d is not None:
          assert os.path.dirname(d["rescale"]) == compat, d["rescale"]
          sc = np.asarray(d.get("scales", []), dtype=float)
          res["dropin"] = {"frames": args.frames, "cold_s": d["times"][0], "warm_s": d["times"][1], "process_wall_s": d["process_wall_s"],
                         "frames_per_s_cold": args.frames / d["times"][0], "frames_per_s_warm": args.frames / d["times"][1],
                         "rescale_module": os.path.relpath(d["rescale"], ROOT),
                         "median_scale_over_true_step": float(np.median(sc[20:] / steps[20:20 + sc[20:].shape[0]])) if sc.shape[0] > 40 else None}
        if args.ref_frames > 0:
            n = min(args.ref_frames, args.frames)
            r = run_arm(ref_src, "", listing(n, "ref_00.txt"), ".ref", 1, tmp)
            assert os.path.dirname(r["rescale"]) == os.path.abspath(ref_src), r["rescale"]
            rs = np.asarray(r.get("scales", []), dtype=float)
            res["reference"] = {"frames": n, "wall_s": r["times"][0], "frames_per_s": n / r["times"][0], "process_wall_s": r["process_wall_s"],
                                "rescale_module": r["rescale"],
                                "median_scale_over_true_step": float(np.median(rs[20:] / steps[20:20 + rs[20:].shape[0]])) if rs.shape[0] > 40 else None}
            if d is not None:
                res["speedup_warm"] = res["dropin"]["frames_per_s_warm"] / res["reference"]["frames_per_s"]
                res["speedup_cold"] = res["dropin"]["frames_per_s_cold"] / res["reference"]["frames_per_s"]
    res["note"] = ("whole-program wall clock of main.py (image decoding, AKAZE, LK, findEssentialMat, recoverPose, estimator, result files); the "
                   "front-end is the same host code in both arms, the arms differ in the module set `import rescale` resolves to")
    print(json.dumps(res))


if __name__ == "__main__":
    main()
