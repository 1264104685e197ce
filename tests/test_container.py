"""Packed frame container (N2) and the reader of the reference's pickled hand-off file: host-side, no GPU."""
import os

import numpy as np
import pytest

from mvoscalerecovery_b200 import container as C


def _ragged(rng, F=7):
    sizes = [0, 5, 130, 1, 64, 0, 33][:F]
    f3 = [rng.standard_normal((n, 3)).astype(np.float32).astype(np.float64) for n in sizes]
    f2 = [rng.uniform(0, 1241, (n, 2)).astype(np.float32).astype(np.float64) for n in sizes]
    mot = [np.hstack([np.eye(3), rng.standard_normal((3, 1))]).reshape(-1) for _ in sizes]
    mv = [bool(b) for b in rng.random(len(sizes)) < 0.8]
    return mot, f3, f2, mv


def test_roundtrip_and_layout(tmp_path):
    rng = np.random.default_rng(1)
    mot, f3, f2, mv = _ragged(rng)
    seq = C.pack_sequence(mot, f3, f2, mv)
    p = str(tmp_path / "s.mvosr")
    C.save_packed(p, seq)
    raw = open(p, "rb").read()
    assert raw[:8] == b"MVOSRPK1" and len(raw) % 64 == 0
    for mm in (True, False):
        got = C.load_packed(p, mmap=mm)
        for k, v in seq.items():
            assert got[k].dtype == v.dtype and np.array_equal(got[k], v), k
    back = C.unpack_sequence(C.load_packed(p))
    assert back["move_flags"] == mv
    for a, b in zip(back["feature3ds"], f3):
        assert np.array_equal(a, b)
    for a, b in zip(back["feature2ds"], f2):
        assert np.array_equal(a, b)
    for a, b in zip(back["motions"], mot):
        assert np.array_equal(a, b)
    assert list(np.diff(seq["offsets"])) == [a.shape[0] for a in f3]


def test_reads_the_reference_hand_off_file(tmp_path):
    """np.save of the dict of ragged lists, as src/main.py:149-154 writes it; file name rule of src/main_offline.py:37."""
    rng = np.random.default_rng(2)
    mot, f3, f2, mv = _ragged(rng)
    path = str(tmp_path / "00_result.npy.tag1")
    np.save(path, {"motions": mot, "feature3ds": f3, "feature2ds": f2, "move_flags": mv})     # numpy appends ".npy"
    seq = C.load_reference_npy(path + ".npy")
    want = C.pack_sequence(mot, f3, f2, mv)
    for k in want:
        assert np.array_equal(seq[k], want[k]), k
    assert C.result_prefix("result/00_result.npy.tag1.npy") == "00_result_"


def test_empty_and_bad_files(tmp_path):
    seq = C.pack_sequence([], [], [], [])
    p = str(tmp_path / "e.mvosr")
    C.save_packed(p, seq)
    got = C.load_packed(p)
    assert got["offsets"].tolist() == [0] and got["x"].shape == (0,)
    bad = str(tmp_path / "bad.mvosr")
    open(bad, "wb").write(b"NOTMVOSR" + b"\0" * 100)
    with pytest.raises(ValueError, match="bad magic"):
        C.load_packed(bad)
    rng = np.random.default_rng(3)
    mot, f3, f2, mv = _ragged(rng)
    C.save_packed(p, C.pack_sequence(mot, f3, f2, mv))
    with open(p, "r+b") as fh:
        fh.truncate(os.path.getsize(p) - 64)
    with pytest.raises(ValueError, match="truncated"):
        C.load_packed(p)
    with pytest.raises(ValueError, match="disagree"):
        C.pack_sequence(mot, f3, [a[:-1] if a.shape[0] else a for a in f2], mv)


@pytest.mark.gpu
def test_offline_driver_reproduces_reference_files(golden, tmp_path, monkeypatch):
    """mvoscalerecovery_b200.offline on the reference's hand-off file: scales.txt == the reference's per-frame scales,
    path.txt == get_path of the motions with those scales (src/main_offline.py:90-119)."""
    from mvoscalerecovery_b200 import offline
    g = golden
    F = g.n_frames
    rng = np.random.default_rng(4)
    mot = [np.hstack([np.eye(3), np.array([[0.01 * rng.standard_normal()], [0.0], [1.0]])]).reshape(-1) for _ in range(F)]
    data = {"motions": mot, "feature3ds": [g.f3(f).astype(np.float64) for f in range(F)],
            "feature2ds": [g.f2(f).astype(np.float64) for f in range(F)], "move_flags": [bool(b) for b in g.z["move_flags"]]}
    monkeypatch.chdir(tmp_path)
    os.makedirs("result")
    np.save("result/07_result.npy.t", data)
    seq = C.load_reference_npy("result/07_result.npy.t.npy")
    res = offline.recover_sequence(seq, absolute_reference=1.7, window_size=5, seed=g.seed)
    np.testing.assert_allclose(res["scales"], g.z["scales"], rtol=1e-9, atol=1e-12)
    poses = np.zeros((F + 1, 12)); cur = np.eye(4); poses[0] = cur[:3].reshape(-1)
    for i in range(F):
        m = np.eye(4); m[:3] = mot[i].reshape(3, 4); m[:3, 3] *= g.z["scales"][i]
        cur = cur @ m; poses[i + 1] = cur[:3].reshape(-1)
    np.testing.assert_allclose(res["poses"], poses, rtol=1e-9, atol=1e-9)
    # the command line: same file names as the reference
    C.save_packed("seq07.mvosr", seq)
    offline.main(["x", "result/07_result.npy.t.npy", ".t"])
    offline.main(["x", "seq07.mvosr", ".t"])
    a = np.loadtxt("evaluate_result/07_result_scales.txt.t"); b = np.loadtxt("evaluate_result/seq07_scales.txt.t")
    assert a.shape == (F,) and np.array_equal(a, b) and np.loadtxt("evaluate_result/07_result_path.txt.t").shape == (F + 1, 12)


@pytest.mark.gpu
def test_gpu_path_on_the_reference_mains_own_hand_off(tmp_path, monkeypatch):
    """BASELINE configs[0] in small: the hand-off written by the reference's unmodified src/main.py on rendered frames (AKAZE +
    LK + findEssentialMat + recoverPose features), through the batched GPU path == the reference's estimator over the same
    hand-off (tests/golden/make_main_golden.py): filtered scales to 1e-9, survivor / triangle / inlier counts exact."""
    import torch
    from mvoscalerecovery_b200 import offline
    from mvoscalerecovery_b200.batch import ScaleRecovery, stats_to_numpy
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "main_c1.npz"))
    seq = {k: z[k] for k in ("offsets", "move_flags", "motions", "x", "y", "z", "u", "v")}
    ref_h = float(z["absolute_reference"])
    res = offline.recover_sequence(seq, absolute_reference=ref_h, window_size=5, seed=int(z["seed"]))
    np.testing.assert_allclose(res["scales"], z["scales"], rtol=1e-9, atol=1e-12)
    called = z["called"]
    np.testing.assert_allclose(res["raw_scale"][called], z["raw_scale"][called], rtol=1e-9)
    eng = ScaleRecovery(absolute_reference=ref_h)
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(eng.device)
    off = seq["offsets"].astype(np.int32)
    r = eng.scale_frames(t(off, np.int32), *[t(seq[k], np.float32) for k in "xyzuv"], int(np.max(np.diff(off))), seed=int(z["seed"]), stats=True)
    st = stats_to_numpy(r["stats"])
    assert np.array_equal(st["n_kept"][called], z["n_kept"][called]) and np.array_equal(st["n_tri"][called], z["n_tri"][called])
    assert np.array_equal(st["best_ic"][called], z["best_ic"][called])
    # the same through the packed container and the command line
    monkeypatch.chdir(tmp_path)
    from mvoscalerecovery_b200 import container as C
    C.save_packed("c1.mvosr", seq)
    import mvoscalerecovery_b200.compat.param as param
    assert param.camera_h == ref_h
    out = offline.main(["x", "c1.mvosr", ".t"])
    assert out["poses"].shape == (called.shape[0] + 1, 12)
