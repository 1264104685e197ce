"""bucket(features, bucket_size=30, density=2) of the reference's detector (src/detector.py:65-95; FeatureDetector.bucket, :18-47)
on the GPU, numpy in / numpy out: the surviving features, cell by cell, as the reference returns them.  The reference shuffles
every cell with numpy's global RNG; here the choice inside a cell is a function of (seed, frame, feature position) -- a Philox
key per feature (include/mvosr.h: mvosr_bucket_frames) -- so that a run can be repeated.  No CPU fallback."""
import numpy as np


def bucket(features, bucket_size=30, density=2, seed=0, frame=0):
    import torch
    from ._gpu import engine
    eng = engine()
    f = np.ascontiguousarray(features, dtype=np.float32).reshape(-1, 2)
    n = f.shape[0]
    if n > 4096:
        raise ValueError("bucket: %d features in one frame; the kernel's key layout holds 4096" % n)
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(eng.device)
    out = eng.bucket_frames(t([0, n], np.int32), t(f[:, 0], np.float32), t(f[:, 1], np.float32), bucket_size=bucket_size, density=density,
                            seed=seed, frame_index=t([frame], np.int32))
    if int(out["status"].cpu()[0]):
        raise ValueError("bucket: feature coordinates must be finite, non-negative pixels")
    k = int(out["n_out"].cpu()[0])
    return f[out["index"][:k].cpu().numpy()]
