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
        import thirdparty.Ransac.ransac as ransac
        assert os.path.dirname(os.path.abspath(rescale.__file__)) == COMPAT
        yield dict(rescale=rescale, graph=graph, ern=estimate_road_norm, sc=scale_calculator, reconstruct=reconstruct,
                   param=param, ransac=ransac)
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
    model, ic = e.get_pitch_ransac(pts, 20, 0.05)
    assert model.shape == (4,) and 3 <= ic <= 40
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
