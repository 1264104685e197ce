"""mvoscalerecovery_b200/compat: the drop-in module set (same names as the reference's src/*.py).  CPU tests of the
API surface against outputs of the reference's own functions (tests/golden/compat_api.npz, written by
tests/golden/make_compat_golden.py); the GPU test drives rescale.ScaleEstimator exactly as src/main_offline.py does."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMPAT = os.path.join(ROOT, "mvoscalerecovery_b200", "compat")


@pytest.fixture(scope="module")
def compat():
    sys.path.insert(0, COMPAT)
    try:
        import rescale, graph, estimate_road_norm, scale_calculator, reconstruct, param       # noqa: E401
        import calculate_height_pitch, triangle_batch                                          # noqa: E401
        import thirdparty.Ransac.ransac as ransac
        assert os.path.dirname(os.path.abspath(rescale.__file__)) == COMPAT
        yield dict(rescale=rescale, graph=graph, ern=estimate_road_norm, sc=scale_calculator, reconstruct=reconstruct,
                   param=param, ransac=ransac, chp=calculate_height_pitch, tb=triangle_batch)
    finally:
        sys.path.remove(COMPAT)


@pytest.fixture(scope="module")
def ref():
    return np.load(os.path.join(ROOT, "tests", "golden", "compat_api.npz"))


def test_constructor_and_attributes_mirror_reference(compat):
    est = compat["rescale"].ScaleEstimator(absolute_reference=1.7, window_size=5)
    for name, val in (("absolute_reference", 1.7), ("camera_pitch", 0), ("scale", 1), ("inliers", None), ("window_size", 5), ("vanish", 185)):
        assert getattr(est, name) == val
    assert len(est.scale_queue) == 0 and est.img_w == compat["param"].img_w and est.img_h == compat["param"].img_h
    assert est.initial_estimation(np.array([0.0, 0.1, 0.99])) == 0
    assert isinstance(est.sc, compat["sc"].ScaleEstimator) and isinstance(est.gc, compat["graph"].GraphChecker)
    old = compat["sc"].ScaleEstimator(1.7)
    assert old.window_size == 6 and old.vanish == 185 and old.focus == 718
    np.testing.assert_allclose(old.initial_estimation(np.array([0.0, 0.1, 0.99])), np.degrees(np.arcsin(0.1)))
    assert compat["param"].camera_h == 1.75 and compat["param"].minimum_feature_for_scale == 100


def test_graph_helpers_equal_reference(compat, ref):
    g = compat["graph"]
    assert np.array_equal(g.triangle([[3, 1], [2, 2], [2, 2], [0, 4]]), ref["tp"])
    np.testing.assert_allclose(g.get_probability([0, 1, 2], [2, 1, 1], ref["tp"]), ref["prob_012_211"], rtol=1e-15)
    keep = g.GraphChecker([[3, 1], [2, 2], [2, 2], [0, 4]]).find_inliers(ref["f3"], ref["f2"], ref["tri"])
    assert np.array_equal(keep, ref["keep"])
    assert g.check_triangle([0, 1, 2], [2, 1, 1]) == 5 and np.array_equal(g.bool2id(np.array([True, False, True])), [0, 2])


def test_estimator_helpers_equal_reference(compat, ref):
    est = compat["rescale"].ScaleEstimator(1.7, 5)
    assert np.array_equal(est.find_outliers(ref["f3"], ref["f2"], ref["tri"]), ref["outliers"])
    ids, hl = est.flat_selection(ref["f3"], ref["tri"])
    assert np.array_equal(np.asarray(ids), ref["flat_ids"])
    np.testing.assert_allclose(hl, ref["flat_heights"], rtol=1e-9)
    np.testing.assert_allclose(est.height_level, ref["height_level"], rtol=1e-9)
    got = np.array([est.check_triangle([0., 1., 2.], [2., 1., 1.]), est.check_triangle([3., 1., 2.], [1., 2., 3.])])
    assert np.array_equal(got, ref["check_triangle"])
    f3 = ref["f3"].copy()
    est.flat_selection(f3, ref["tri"]); est.find_outliers(f3, ref["f2"], ref["tri"])
    assert np.array_equal(f3, ref["f3"])                     # inputs are never modified


def test_road_norm_helpers_equal_reference(compat, ref):
    e = compat["ern"]
    pts = ref["f3"][:40]
    m = e.estimate(pts[:3])
    np.testing.assert_allclose(m * np.sign(m[1]), ref["plane"], rtol=1e-9, atol=1e-12)
    assert np.array_equal(e.get_inliers(m, pts, 0.05), ref["inl"])
    np.testing.assert_allclose(e.get_pitch(ref["ts"]), ref["pitch"], rtol=1e-12)
    np.testing.assert_allclose(e.get_pitch_svd(ref["ts"]), ref["pitch_svd"], rtol=1e-9)
    np.testing.assert_allclose(np.asarray(e.get_norm_svd(ref["ts"])).reshape(-1), ref["norm_svd"], rtol=1e-9)
    assert bool(e.is_inlier(m, pts[0], 1e-6)) and e.augment(pts[:2]).shape == (2, 4)
    for name in ("run_ransac", "random", "np", "math", "sys", "Delaunay"):       # names the reference re-exports via import *
        assert hasattr(e, name), name


def test_run_ransac_bookkeeping(compat):
    """First strictly larger count is kept; stops at the first count above the goal (ransac.py:9-22)."""
    rr = compat["ransac"].run_ransac
    calls = []

    def estimate(s):
        calls.append(1)
        return len(calls)
    counts = {1: 3, 2: 5, 3: 5, 4: 9, 5: 2}
    model, ic = rr(list(range(10)), estimate, lambda m, x: x < counts.get(m, 0), 3, 8, 20)
    assert (model, ic) == (4, 9) and len(calls) == 4


@pytest.fixture(scope="module")
def sref():
    """Outputs of the reference's older estimator and batch scripts (tests/golden/make_script_golden.py)."""
    return np.load(os.path.join(ROOT, "tests", "golden", "scripts.npz"))


def _ragged(flat, lens):
    out, a = [], 0
    for n in lens:
        out.append(list(flat[a:a + n])); a += n
    return out


def test_older_estimator_host_analysis_equals_reference(compat, sref):
    """Histogram / mode / skewness analysis and the mesh graphs of src/scale_calculator.py (host code)."""
    e = compat["sc"].ScaleEstimator(1.7)
    sel = sref["sc_f3_1"][sref["sc_by_tri"]]
    dis, bins = np.histogram(sel[:, 1], bins=np.arange(170) * 0.1)
    assert np.array_equal(dis, sref["sc_hist"])
    assert np.array_equal(e.check_reverse_mode(dis), sref["sc_reverse_mode"])
    modes = e.check_mode(dis.copy(), bins)
    assert [np.asarray(g).size for g in modes] == list(sref["sc_modes_len"])
    assert np.array_equal(np.concatenate([np.asarray(g, float).reshape(-1) for g in modes]), sref["sc_modes_flat"])
    assert e.check_skewness(sel[:, 1]) == sref["sc_skew_p1"] and e.check_skewness(sel[:, 1], method="p2") == sref["sc_skew_p2"]
    e.height_level = float(sref["sc_by_tri_level"])
    assert e.road_model_calculation_static(sel.copy())[0] == sref["sc_static"]
    assert e.road_model_calculation_static_tri(sref["sc_static_tri_in"])[0] == sref["sc_static_tri"]
    tri = sref["sc_tri"]
    g = e.triangle2graph(tri[:50]); rg = e.triangle2region_graph(tri[:50])
    assert [len(x) for x in g] == list(sref["sc_graph_len"]) and [y for x in g for y in x] == list(sref["sc_graph_flat"])
    assert [len(x) for x in rg] == list(sref["sc_rgraph_len"]) and [y for x in rg for y in x] == list(sref["sc_rgraph_flat"])
    f2 = sref["frame2"][:, :2]
    assert np.array_equal(e.find_reliability_by_graph(sref["sc_f3_1"], f2, tri), sref["sc_reliability"])
    # remove_single / compare / check_depth / scale_filtering: plain semantics
    assert (e.compare(1.0, 1.05), e.compare(1.0, 1.2), e.compare(1.2, 1.0)) == (0, -1, 1)
    assert e.check_depth([2.0, 1.0], [3.0, 1.0]) and not e.check_depth([2.0, 1.0], [1.0, 3.0])
    f = np.array([[0, 0.05, 1], [0, 0.55, 1], [0, 0.56, 1], [0, 0.57, 1.0]])
    d, b = np.histogram(f[:, 1], bins=np.arange(170) * 0.1)
    assert e.remove_single(f, d, b).shape[0] == 3
    assert [e.scale_filtering(x) for x in (1.0, 3.0, 2.0)] == [1.0, 2.0, 2.0]


def test_graph_grow_regions(compat):
    """GraphGrow.process (graph.py:84-107): largest similarity-connected region holding a flat, low triangle."""
    gg = compat["graph"].GraphGrow(threshold_angle=8)
    # a strip of 8 triangles sharing consecutive edges: 0-4 flat and similar, 5 steep (breaks the chain), 6-7 flat
    tri = np.array([[i, i + 1, i + 2] for i in range(8)])
    angles = np.array([-88.0, -87.0, -86.5, -88.0, -87.5, -40.0, -86.0, -88.0])
    heights = np.array([1.7, 1.72, 1.69, 1.71, 1.7, 1.2, 1.65, 1.66])
    got = gg.process(tri, heights, angles)
    hinv = 1 / heights
    seeds = np.nonzero((angles < -85) & (hinv < np.median(hinv[angles < -80])))[0]
    assert sorted(got) == [0, 1, 2, 3, 4] and got[0] == seeds[0]
    assert gg.graph[3] == [2, 4] and abs(gg.threshold_height - 0.4 * np.median(1 / heights)) < 1e-15
    assert gg.process(tri, heights, np.full(8, -30.0)) == []


def test_product_path_fails_loudly_without_a_gpu(compat):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    est = compat["rescale"].ScaleEstimator(1.7, 5)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        est.scale_calculation(np.zeros((200, 3)), np.zeros((200, 2)))


@pytest.mark.gpu
def test_drop_in_estimator_reproduces_reference_sequence(compat, golden):
    """The loop of src/main_offline.py:57-88 around the drop-in ScaleEstimator == the reference's per-frame outputs."""
    g = golden
    est = compat["rescale"].ScaleEstimator(absolute_reference=1.7, window_size=5)
    est.seed = g.seed
    scales = [0]
    for f in range(g.n_frames):
        if not g.z["move_flags"][f]:
            scales.append(0)
            est.frame_index = f + 1
            continue
        f3, f2 = g.f3(f).astype(np.float64), g.f2(f).astype(np.float64)
        if f3.shape[0] > compat["param"].minimum_feature_for_scale:
            est.frame_index = f
            est.initial_estimation(np.zeros(3))
            s, std = est.scale_calculation(f3, f2)
            assert std == 1
            scales.append(s)
        else:
            scales.append(scales[-1])
    np.testing.assert_allclose(np.asarray(scales[1:], float), g.z["scales"], rtol=1e-9, atol=1e-12)
    # feature_selection returns what the reference returns for one frame
    f = next(i for i in range(g.n_frames) if g.called(i))
    est2 = compat["rescale"].ScaleEstimator(1.7, 5)
    pts, heights = est2.feature_selection(g.f3(f).astype(np.float64), g.f2(f).astype(np.float64))
    fl = g.get(f, "flags")
    np.testing.assert_allclose(heights, g.get(f, "heights")[(fl & 1) != 0], rtol=1e-7)
    roi = g.f3(f).astype(np.float64)[g.f2(f)[:, 1] > 185]
    kept = roi[g.get(f, "keep")] if g.get(f, "keep").sum() > 10 else roi
    assert np.array_equal(pts, kept[g.get(f, "data_id")])


@pytest.mark.gpu
def test_older_estimator_reproduces_reference(compat, sref):
    """src/scale_calculator.py::ScaleEstimator.scale_calculation frame by frame (Delaunay, votes and planes on the GPU)."""
    FX, CX, CY = 718.856, 607.1928, 185.2157
    est = compat["sc"].ScaleEstimator(1.7, window_size=5)
    for f in range(int(sref["n_frames"])):
        f3 = sref["sc_f3_%d" % f].copy()
        f2 = sref["frame%d" % (f + 1)][:, :2].copy()
        s, std = est.scale_calculation(f3, f2)
        np.testing.assert_allclose(s, sref["sc_scales"][f], rtol=1e-12)
        np.testing.assert_allclose(est.height_level, sref["sc_levels"][f], rtol=1e-9)
        assert std == sref["sc_stds"][f] and est.flat_feature.shape[0] == sref["sc_nsel"][f]
        if f == 0:
            np.testing.assert_allclose(f3, sref["sc_remapped0"], rtol=1e-15)           # remapped in place, like the reference
            np.testing.assert_allclose(est.flat_feature, sref["sc_flat0"], rtol=1e-15)
    e2 = compat["sc"].ScaleEstimator(1.7)
    f3, f2, tri = sref["sc_f3_1"], sref["frame2"][:, :2], sref["sc_tri"]
    assert np.array_equal(e2.find_outliers(f3, f2, tri), sref["sc_find_outliers"])
    assert np.array_equal(e2.feature_selection_by_tri(f3, tri), sref["sc_by_tri"])
    np.testing.assert_allclose(e2.height_level, sref["sc_by_tri_level"], rtol=1e-12)
    assert np.array_equal(e2.feature_selection_by_tri_graph(f3, tri), sref["sc_by_tri_graph"])
    # Delaunay helper == canonicalised Qhull
    import _gpu
    assert np.array_equal(_gpu.delaunay(f2), tri)
    # Reconstruct on the same mesh
    rec = compat["reconstruct"].Reconstruct()
    tm = rec.triangle_model(f3, tri)
    n = np.stack([np.linalg.inv(f3[t]) @ np.ones(3) for t in tri])
    ln = np.linalg.norm(n, axis=1); sg = np.where(n[:, 1] < 0, -1.0, 1.0)
    np.testing.assert_allclose(tm, np.hstack([n / ln[:, None] * sg[:, None], (sg / ln)[:, None]]), rtol=1e-7, atol=1e-11)
    np.testing.assert_allclose(rec.find_outliers(f3, f2, tri), compat["rescale"].ScaleEstimator(1.7).find_outliers(f3, f2, tri))


@pytest.mark.gpu
def test_batch_scripts_reproduce_reference(compat, sref, tmp_path, monkeypatch):
    """calculate_height_pitch.py / triangle_batch.py: same command line, same files, same numbers as the Python-2 scripts."""
    chp, tb = compat["chp"], compat["tb"]
    F = int(sref["n_frames"])
    feat = tmp_path / "feat"
    feat.mkdir()
    for k in range(1, F + 1):
        np.savetxt(feat / ("%d.txt" % k), sref["frame%d" % k], fmt="%.9g")
    (tmp_path / "images.txt").write_text("\n".join("img%d.png" % i for i in range(F + 2)) + "\n")
    np.savetxt(tmp_path / "motions.txt", sref["motions"])
    monkeypatch.chdir(tmp_path)
    monkeypatch.setattr(chp, "ransac_seed", int(sref["seed"]))
    res = chp.main(["x", str(tmp_path / "images.txt"), str(feat) + os.sep, str(tmp_path / "motions.txt"), str(tmp_path / "motions.txt")])
    for name, key in zip(chp.OUTPUT_FILES, ("chp_heights", "chp_h_means", "chp_h_stds", "chp_h_t_means", "chp_pitches", "chp_inlier_numbers")):
        got = np.loadtxt(tmp_path / name)
        np.testing.assert_allclose(got, sref[key], rtol=1e-7, err_msg=name)
    assert np.array_equal(res["inlier_numbers"], sref["chp_inlier_numbers"])
    heights = tb.main(["x", str(tmp_path / "images.txt"), str(feat) + os.sep])
    np.testing.assert_allclose(heights, sref["tb_heights"], rtol=1e-9)


@pytest.mark.gpu
def test_get_pitch_ransac_runs_on_the_gpu_and_repeats(compat, ref):
    from oracle import pipeline as P
    e = compat["ern"]
    pts = ref["f3"][:200]
    e.ransac_seed, e.ransac_calls = 5, 7
    m, ic = e.get_pitch_ransac(pts, 50, 0.05)
    r = P.ransac_plane(pts, 5, 7, max_iterations=50, thr=0.05)
    assert ic == r["ic"] and e.ransac_calls == 8
    np.testing.assert_allclose(m, np.asarray(r["model"]) * np.sign(r["model"][1]), rtol=1e-7, atol=1e-11)
    h, pitch, inl = compat["sc"].ScaleEstimator(1.7).road_model_calculation_ransac(pts)
    assert inl.shape[1] == 3 and 0 < inl.shape[0] <= 200 and np.isfinite(h) and abs(pitch) < np.pi / 2


def test_drop_in_directory_coexists_with_the_reference_tree():
    """src/main.py imports thirdparty.MonocularVO.visual_odometry (its own VO front-end) AND rescale -> thirdparty.Ransac.ransac:
    with the drop-in directory first on the path, `thirdparty` must stay a namespace package so that MonocularVO still resolves
    to the reference tree while Ransac, rescale, ... resolve here.  Needs the reference tree (build container only)."""
    import subprocess
    ref_src = "/root/reference/src"
    if not os.path.isfile(os.path.join(ref_src, "main.py")):
        pytest.skip("reference tree not present")
    assert not os.path.exists(os.path.join(COMPAT, "thirdparty", "__init__.py"))
    code = (
        "import sys, types, numpy as np\n"
        "sys.modules['matplotlib'] = types.ModuleType('matplotlib'); sys.modules['matplotlib.pyplot'] = types.ModuleType('matplotlib.pyplot')\n"
        "sys.path[:0] = [%r, %r]\n"
        "import thirdparty.MonocularVO.visual_odometry as vo, thirdparty.Ransac.ransac as rr, rescale, scale_calculator, graph, estimate_road_norm, param\n"
        "print(vo.__file__); print(rr.__file__); print(rescale.__file__); print(param.__file__)\n"
        "est = rescale.ScaleEstimator(absolute_reference=param.camera_h, window_size=5)\n"
        "print(est.initial_estimation(np.zeros(3)), hasattr(vo, 'VisualOdometry'))\n" % (COMPAT, ref_src))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = out.stdout.strip().splitlines()
    assert lines[0].startswith(ref_src) and lines[1].startswith(COMPAT) and lines[2].startswith(COMPAT) and lines[3].startswith(COMPAT)
    assert lines[4] == "0 True"
