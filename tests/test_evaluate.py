"""mvoscalerecovery_b200.evaluate against outputs of the reference's own evaluation scripts (script/evaluate_vo.py,
script/evaluate_scale.py; goldens written by tests/golden/make_script_golden.py).  Host code: no GPU."""
import os

import numpy as np
import pytest

from mvoscalerecovery_b200 import evaluate as E

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(ROOT, "tests", "golden", "scripts.npz"))


def test_kitti_segment_errors_equal_reference(g):
    np.testing.assert_allclose(E.trajectory_distances(g["ev_gt"]), g["ev_dist"], rtol=1e-12)
    errs = E.sequence_errors(g["ev_gt"], g["ev_res"])
    assert errs.shape == g["ev_errors"].shape and errs.shape[0] > 100
    assert np.array_equal(errs[:, [0, 3]], g["ev_errors"][:, [0, 3]])                  # start frames and lengths: exact
    np.testing.assert_allclose(errs, g["ev_errors"], rtol=1e-7, atol=1e-12)
    rot, tra, tra_all = E.average_errors(errs)
    np.testing.assert_allclose(rot, g["ev_rot"], rtol=1e-7)
    np.testing.assert_allclose(tra, g["ev_tra"], rtol=1e-7)
    assert len(tra_all) == len(E.LENGTHS)
    assert abs(E.kitti_translation_error(g["ev_gt"], g["ev_res"]) - np.mean(g["ev_tra"])) < 1e-9
    assert E.kitti_translation_error(g["ev_gt"], g["ev_gt"]) < 1e-12
    assert np.isnan(E.kitti_translation_error(g["ev_gt"][:50], g["ev_res"][:50]))      # too short for any segment


def test_scale_statistics_equal_reference(g):
    gt, re = g["es_gt"], g["es_re"]
    np.testing.assert_allclose(E.patch(gt[:re.shape[0]] - re, 50, 10), g["es_patch"], rtol=1e-9, atol=1e-14)
    assert np.array_equal(E.filter(re, 10), g["es_filter"])
    s = E.scale_errors(gt, re)
    er = np.abs(gt[:re.shape[0]] - re)
    assert s["mean"] == np.mean(er) and s["max"] == np.max(er) and s["within_0.1"] == 1 - np.sum(er > 0.1) / re.shape[0]
    assert len(s["windowed"]) == 11 and E.patch(np.ones(5), 10).shape == (0,)


def test_command_line(g, tmp_path, capsys):
    np.savetxt(tmp_path / "gt.txt", g["ev_gt"]); np.savetxt(tmp_path / "re.txt", g["ev_res"])
    E.main(["x", "vo", str(tmp_path / "gt.txt"), str(tmp_path / "re.txt")])
    out = capsys.readouterr().out.strip().splitlines()[-1].split()
    assert abs(float(out[0]) - np.mean(g["ev_tra"])) < 1e-9
    with pytest.raises(SystemExit):
        E.main(["x", "bogus"])
