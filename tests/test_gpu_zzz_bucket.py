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
