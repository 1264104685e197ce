"""mvosr_bucket_frames (feature bucketing, src/detector.py:65-95) on the GPU through the C ABI, against the oracle index
for index: the golden feature sets of the reference's own bucket() run, a frame_index map, and the edge frames (empty, one
feature, 4096 and 4097 features, a negative coordinate, a NaN).  Runs last among the GPU tests (file name): newest kernel, written
and CPU-verified (pthread emulation of the kernel source, tests/test_bucket.py) after the round's GPU minutes were spent."""
import numpy as np
import pytest

from test_bucket import _batch, _check_against_oracle, golden          # noqa: F401  (golden is a fixture)

pytestmark = pytest.mark.gpu


def test_bucket_frames_on_the_gpu(engine, golden):
    import torch
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(engine.device)
    for sets, bs, dens in (((0, 1, 4), 30, 2), ((2,), 50, 1), ((3, 4, 0), 20, 3)):
        off, u, v = _batch(golden, sets)
        fidx = (np.arange(len(sets)) * 7 + 2).astype(np.int32)
        out = engine.bucket_frames(t(off), t(u), t(v), bucket_size=bs, density=dens, seed=2**35 + 11, frame_index=t(fidx), seq_id=4)
        torch.cuda.synchronize()
        _check_against_oracle(off, u, v, bs, dens, 2**35 + 11, fidx, 4, out["index"].cpu().numpy(), out["n_out"].cpu().numpy(), out["status"].cpu().numpy())
    rng = np.random.default_rng(0)
    lens = [0, 1, 4096, 4097, 30, 30, 5]
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    u = rng.uniform(0, 1241, off[-1]).astype(np.float32); v = rng.uniform(0, 376, off[-1]).astype(np.float32)
    u[off[4] + 3] = -1.0; v[off[5] + 7] = np.nan
    out = engine.bucket_frames(t(off), t(u), t(v), seed=5)
    torch.cuda.synchronize()
    _check_against_oracle(off, u, v, 30, 2, 5, None, 0, out["index"].cpu().numpy(), out["n_out"].cpu().numpy(), out["status"].cpu().numpy())


def test_compat_bucket_returns_what_the_reference_returns(engine, golden):
    """compat.bucketing.bucket: the reference's signature and return type; same cells in the same order with the same number of
    survivors as the reference's own output (tests/golden/bucket.npz), every survivor a member of its cell."""
    from mvoscalerecovery_b200.compat import bucketing
    from test_bucket import _cells
    for k in (0, 2, 4):
        f, kept = golden["f%d" % k], golden["kept%d" % k]
        bs, dens = (int(x) for x in golden["par%d" % k])
        got = bucketing.bucket(f, bs, dens, seed=1, frame=k)
        assert got.dtype == np.float32 and got.shape == kept.shape
        assert _cells(got, bs) == _cells(kept, bs)
        members = {tuple(p) for p in f.tolist()}
        assert all(tuple(p) in members for p in got.tolist())
    with pytest.raises(ValueError):
        bucketing.bucket(np.array([[1.0, -3.0], [2.0, 2.0]]))
