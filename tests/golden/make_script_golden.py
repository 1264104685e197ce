"""Writes tests/golden/scripts.npz: outputs of the REFERENCE's own batch scripts (src/calculate_height_pitch.py,
src/triangle_batch.py) and of its older estimator (src/scale_calculator.py) on small seeded inputs.

Run in the build container only (needs /root/reference):  python tests/golden/make_script_golden.py

The two scripts are Python-2 top-level programs.  They are executed here from an in-memory copy with exactly two mechanical
edits -- ``print x`` statements turned into ``print(x)`` calls and the hard-coded frame count 4541 replaced by the number
of synthetic frames + 1 -- under the harness shims of oracle/ref_harness.py (matplotlib stub, np.float, canonicalised
Delaunay, Philox sampler whose frame counter follows the feature file being read).  Nothing of the reference is written
into the repository; only its numeric outputs are.
"""
import contextlib
import io
import os
import re
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness as H          # noqa: E402

SEED = 4242
N_FRAMES = 6
FX, CX, CY = 718.856, 607.1928, 185.2157


def synth_frames(rng):
    """(u, v, depth) rows of a road plane 1.7 m below the camera, pitched by ~1 degree, 25 % lifted outliers, plus the
    relative motions (unit steps along a slightly descending direction)."""
    frames = []
    for f in range(N_FRAMES):
        n = 260 + 20 * f
        u = rng.uniform(5, 1236, n); v = rng.uniform(200, 372, n)
        z = 1.7 * FX / (v - CY) * (1 + 0.004 * rng.standard_normal(n))
        bad = rng.random(n) < 0.25
        z[bad] *= rng.uniform(0.55, 0.9, bad.sum())
        frames.append(np.stack([u, v, z], 1).astype(np.float32).astype(np.float64))
    motions = np.zeros((N_FRAMES + 2, 12))
    for i in range(motions.shape[0]):
        t = np.array([0.01 * rng.standard_normal(), 0.015 + 0.004 * rng.standard_normal(), 1.0])
        motions[i] = np.hstack([np.eye(3), (t / np.linalg.norm(t))[:, None]]).reshape(-1)
    return frames, motions


def py3_source(path, n_frames):
    src = open(path).read()
    src = re.sub(r"^(\s*)print (.+)$", r"\1print(\2)", src, flags=re.M)
    assert "4541" in src
    return src.replace("4541", str(n_frames + 1))


def run_script(ns, name, argv, workdir, n_frames):
    """Execute the converted script in a fresh namespace; returns captured stdout."""
    import scipy.spatial
    real_dt, real_loadtxt = scipy.spatial.Delaunay, np.loadtxt

    def loadtxt(fname, *a, **k):
        m = re.search(r"(\d+)\.txt$", str(fname))
        if m:
            ns.shim.begin_frame(int(m.group(1)))
        return real_loadtxt(fname, *a, **k)

    class _Anything:
        def __getattr__(self, _):
            return lambda *a, **k: _Anything()
    plt = sys.modules["matplotlib.pyplot"]
    for fn in ("figure", "Polygon", "gca", "savefig", "close", "imshow", "show"):
        setattr(plt, fn, lambda *a, **k: _Anything())
    real_sample = ns.shim.sample

    def sample(pop, k):                      # the line RANSAC of calculate_height_pitch.py draws 2: result only printed
        return real_sample(pop, 3)[:2] if int(k) == 2 else real_sample(pop, k)
    ns.shim.sample = sample
    buf = io.StringIO()
    old_argv, old_cwd = sys.argv, os.getcwd()
    try:
        scipy.spatial.Delaunay = ns.CanonDelaunay
        np.loadtxt = loadtxt
        sys.argv = argv
        os.chdir(workdir)
        with contextlib.redirect_stdout(buf):
            exec(compile(py3_source(os.path.join(H.REF_SRC, name), n_frames), name, "exec"), {"__name__": "__main__"})
    finally:
        scipy.spatial.Delaunay, np.loadtxt, sys.argv = real_dt, real_loadtxt, old_argv
        ns.shim.sample = real_sample
        os.chdir(old_cwd)
    return buf.getvalue()


def main():
    ns = H.load_reference(seed=SEED)
    rng = np.random.default_rng(SEED)
    frames, motions = synth_frames(rng)
    out = dict(seed=np.int64(SEED), n_frames=np.int64(N_FRAMES), motions=motions)
    for f, a in enumerate(frames):
        out["frame%d" % (f + 1)] = a
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "feat"))
        os.makedirs(os.path.join(tmp, "result"))
        for f, a in enumerate(frames):
            np.savetxt(os.path.join(tmp, "feat", "%d.txt" % (f + 1)), a, fmt="%.9g")
        import cv2
        for i in range(N_FRAMES + 2):                      # the scripts imread every listed image (only to draw on it)
            cv2.imwrite(os.path.join(tmp, "img%d.png" % i), np.zeros((8, 8, 3), np.uint8))
        open(os.path.join(tmp, "images.txt"), "w").write("\n".join("img%d.png" % i for i in range(N_FRAMES + 2)) + "\n")
        np.savetxt(os.path.join(tmp, "motions.txt"), motions)
        np.savetxt(os.path.join(tmp, "poses.txt"), motions)
        feat = os.path.join(tmp, "feat") + os.sep
        run_script(ns, "calculate_height_pitch.py", ["x", os.path.join(tmp, "images.txt"), feat, os.path.join(tmp, "motions.txt"),
                                                    os.path.join(tmp, "poses.txt")], tmp, N_FRAMES)
        for key, name in (("chp_heights", "result_heights_line_ransac.txt"), ("chp_h_means", "refined_camera_height_means.txt"),
                          ("chp_h_stds", "refined_camera_height_stds.txt"), ("chp_h_t_means", "refined_camera_height_t_means.txt"),
                          ("chp_pitches", "refined_pitch.txt"), ("chp_inlier_numbers", "inlier_numbers.txt")):
            out[key] = np.loadtxt(os.path.join(tmp, name))
        txt = run_script(ns, "triangle_batch.py", ["x", os.path.join(tmp, "images.txt"), feat], tmp, N_FRAMES)
        out["tb_heights"] = np.array([float(x) for x in txt.split()])
        assert out["tb_heights"].shape[0] == N_FRAMES and out["chp_heights"].shape[0] == N_FRAMES

    # ---- the older estimator (src/scale_calculator.py) on frames given as (feature3d, feature2d)
    with contextlib.redirect_stdout(io.StringIO()):
        import scale_calculator as sc_ref                   # the reference's module (REF_SRC is first on sys.path)
        assert os.path.dirname(os.path.abspath(sc_ref.__file__)) == os.path.abspath(H.REF_SRC)
        # the vote rule of the older estimator is not symmetric in the triangle's vertices (scale_calculator.py:110-118 flags
        # vertices 0 and 1 for the 0-2 edge), so its result depends on Qhull's vertex order inside a simplex: canonicalised,
        # like rescale.Delaunay in the harness
        sc_ref.Delaunay = ns.CanonDelaunay
        est = sc_ref.ScaleEstimator(1.7, window_size=5)
        scales, stds, levels, nsel = [], [], [], []
        for f, a in enumerate(frames):
            z = a[:, 2]
            f3 = np.stack([(a[:, 0] - CX) * z / FX, (a[:, 1] - CY) * z / FX, z], 1).astype(np.float32).astype(np.float64)
            f2 = a[:, :2].copy()
            out["sc_f3_%d" % f] = f3.copy()
            s, std = est.scale_calculation(f3, f2)
            scales.append(s); stds.append(std); levels.append(est.height_level)
            nsel.append(-1 if est.flat_feature is None else est.flat_feature.shape[0])
            if f == 0:
                out["sc_remapped0"] = f3.copy()             # feature_remap works in place
                out["sc_flat0"] = est.flat_feature.copy()
        out.update(sc_scales=np.array(scales), sc_stds=np.array(stds, float), sc_levels=np.array(levels), sc_nsel=np.array(nsel))
        # pieces with caller-supplied triangles
        from scipy.spatial import Delaunay
        f3 = out["sc_f3_1"]; f2 = frames[1][:, :2]
        tri = H.canonicalise(Delaunay(f2).simplices)
        e2 = sc_ref.ScaleEstimator(1.7)
        out["sc_tri"] = tri
        out["sc_find_outliers"] = e2.find_outliers(f3, f2, tri)
        out["sc_reliability"] = e2.find_reliability_by_graph(f3, f2, tri)
        out["sc_by_tri"] = e2.feature_selection_by_tri(f3, tri); out["sc_by_tri_level"] = np.float64(e2.height_level)
        out["sc_by_tri_graph"] = e2.feature_selection_by_tri_graph(f3, tri)
        sel = f3[out["sc_by_tri"]]
        out["sc_static"] = np.array(e2.road_model_calculation_static(sel.copy())[0])
        hts = 1.0 / np.abs(rng.normal(0.6, 0.08, 300))
        out["sc_static_tri_in"] = hts
        out["sc_static_tri"] = np.array(e2.road_model_calculation_static_tri(hts)[0])
        dis, bins = np.histogram(sel[:, 1], bins=np.array(range(0, 170)) * 0.1)
        out["sc_hist"] = dis.copy()
        out["sc_reverse_mode"] = e2.check_reverse_mode(dis)
        modes = e2.check_mode(dis.copy(), bins)
        out["sc_modes_flat"] = np.concatenate([np.asarray(g, float).reshape(-1) for g in modes]) if modes else np.zeros(0)
        out["sc_modes_len"] = np.array([np.asarray(g).size for g in modes])
        out["sc_skew_p1"] = np.float64(e2.check_skewness(sel[:, 1])); out["sc_skew_p2"] = np.float64(e2.check_skewness(sel[:, 1], method='p2'))
        g = e2.triangle2graph(tri[:50]); rg = e2.triangle2region_graph(tri[:50])
        out["sc_graph_flat"] = np.concatenate([np.asarray(x, np.int64) for x in g if len(x)]); out["sc_graph_len"] = np.array([len(x) for x in g])
        out["sc_rgraph_flat"] = np.concatenate([np.asarray(x, np.int64) for x in rg if len(x)]); out["sc_rgraph_len"] = np.array([len(x) for x in rg])
    # ---- the evaluation scripts (script/evaluate_vo.py, script/evaluate_scale.py) on a synthetic trajectory pair
    import importlib.util
    mods = {}
    for name in ("evaluate_vo", "evaluate_scale"):
        spec = importlib.util.spec_from_file_location("ref_" + name, os.path.join(os.path.dirname(H.REF_SRC), "script", name + ".py"))
        mods[name] = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mods[name])
    n = 1500
    def traj(noise):
        cur = np.eye(4); out_p = [cur[:3].reshape(-1).copy()]
        r = np.random.default_rng(99)
        for i in range(n):
            yaw = 0.004 * np.sin(i / 60.0) + noise * 0.0004 * r.standard_normal()
            m = np.eye(4)
            m[:3, :3] = [[np.cos(yaw), 0, np.sin(yaw)], [0, 1, 0], [-np.sin(yaw), 0, np.cos(yaw)]]
            m[:3, 3] = [0.0, 0.0, 0.85 * (1 + noise * 0.03 * r.standard_normal())]
            cur = cur @ m
            out_p.append(cur[:3].reshape(-1).copy())
        return np.array(out_p)
    gt, res = traj(0.0), traj(1.0)
    with contextlib.redirect_stdout(io.StringIO()):
        errs = np.array(mods["evaluate_vo"].calculate_sequence_error(gt, res))
        rot, tra, _ = mods["evaluate_vo"].calculate_ave_errors(errs)
        sgt = 0.8 + 0.3 * np.sin(np.arange(1200) / 40.0); sre = sgt[:1100] + 0.08 * rng.standard_normal(1100)
        out.update(ev_gt=gt, ev_res=res, ev_errors=errs, ev_rot=np.asarray(rot), ev_tra=np.asarray(tra),
                   ev_dist=np.asarray(mods["evaluate_vo"].trajectory_distances(gt)),
                   es_gt=sgt, es_re=sre, es_patch=mods["evaluate_scale"].patch(sgt[:1100] - sre, 50, 10),
                   es_filter=mods["evaluate_scale"].filter(sre, 10))
    # ---- get_path / motion2pose of src/main_offline.py:95-119 (the module reads sys.argv only inside main())
    spec = importlib.util.spec_from_file_location("ref_main_offline", os.path.join(H.REF_SRC, "main_offline.py"))
    mo = importlib.util.module_from_spec(spec)
    with contextlib.redirect_stdout(io.StringIO()):
        spec.loader.exec_module(mo)
    gp_mot = np.tile(np.hstack([np.eye(3), [[0.0], [0.0], [1.0]]]).reshape(-1), (60, 1)) + 0.01 * rng.standard_normal((60, 12))
    gp_sc = rng.uniform(0.0, 1.4, 60)
    out.update(gp_motions=gp_mot.copy(), gp_scales=gp_sc, gp_poses=np.asarray(mo.get_path(gp_mot.copy(), gp_sc)))
    path = os.path.join(ROOT, "tests", "golden", "scripts.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: np.asarray(v).shape for k, v in out.items() if not k.startswith("frame") and not k.startswith("sc_f3")})
    for k in ("chp_heights", "chp_h_means", "chp_inlier_numbers", "tb_heights", "sc_scales", "sc_nsel", "sc_levels"):
        print(k, out[k])


if __name__ == "__main__":
    main()
