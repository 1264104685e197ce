"""Batched counterpart of the reference's offline driver (src/main_offline.py): same input file, same output files,
one GPU pass instead of the per-frame Python loop.

    python -m mvoscalerecovery_b200.offline <name>_result.npy<tag>.npy [tag]        # the reference's pickled hand-off
    python -m mvoscalerecovery_b200.offline sequence.mvosr [tag]                     # the packed container (container.py)

writes ``evaluate_result/<prefix>scales.txt<tag>`` and ``evaluate_result/<prefix>path.txt<tag>`` (prefix rule of :37) exactly as
src/main_offline.py:90-93 does: the filtered per-frame scales (``scales[1:]``) and the N+1 poses of the rescaled
trajectory (get_path, :95-119).  Stages 2-5 run in one launch of the fused frame kernel over the whole sequence, the
driver gating and the temporal state in filter_kernel, the trajectory in integrate_paths_kernel.  No CPU fallback.
"""
from __future__ import annotations

import os
import sys

import numpy as np

from . import container


def recover_sequence(seq: dict, absolute_reference: float, window_size: int = 5, seed: int = 0, seq_id: int = 0, engine=None):
    """seq: arrays in the container's layout.  Returns dict(scales (F,), raw_scale, status, poses (F+1,12), filter10)."""
    import torch
    from .batch import ScaleRecovery
    eng = engine or ScaleRecovery(absolute_reference=float(absolute_reference), window_size=int(window_size))
    dev = eng.device
    t = lambda a, dt: torch.from_numpy(np.array(a, dtype=dt, order="C")).to(dev)      # (a copy: memory-mapped sections are read-only)
    off = np.asarray(seq["offsets"], np.int32)
    F = off.shape[0] - 1
    d = {k: t(seq[k], np.float32) for k in "xyzuv"}
    d_off = t(off, np.int32)
    maxf = int(np.max(np.diff(off))) if F else 0
    r = eng.scale_frames(d_off, d["x"], d["y"], d["z"], d["u"], d["v"], max(maxf, 1), seed=seed, seq_id=seq_id, stats=False)
    nfeat = t(np.diff(off), np.int32)
    seq_off = t(np.array([0, F], np.int32), np.int32)
    out = eng.filter_sequences(seq_off, r["raw_scale"], r["status"], t(seq["move_flags"], np.uint8), nfeat, filter10=True)
    poses = eng.integrate_paths(seq_off, t(seq["motions"], np.float64), out["scale"])
    torch.cuda.synchronize(dev)
    return dict(scales=out["scale"].cpu().numpy(), filter10=out["filter10"].cpu().numpy(), raw_scale=r["raw_scale"].cpu().numpy(),
                status=r["status"].cpu().numpy(), poses=poses.cpu().numpy())


def main(argv=None):
    argv = sys.argv if argv is None else argv
    if len(argv) < 2:
        sys.exit("usage: python -m mvoscalerecovery_b200.offline <result.npy | sequence.mvosr> [tag]")
    data_path = argv[1]
    tag = argv[2] if len(argv) > 2 else '.test_001'
    if data_path.endswith(".mvosr"):
        seq = container.load_packed(data_path)
        prefix = os.path.basename(data_path)[:-len(".mvosr")] + "_"
    else:
        seq = container.load_reference_npy(data_path)
        prefix = container.result_prefix(data_path)
    try:
        import param                                     # the reference's config module when it is on the path
        camera_h = param.camera_h
    except ImportError:
        from .compat import param as _param
        camera_h = _param.camera_h
    res = recover_sequence(seq, absolute_reference=camera_h, window_size=5)
    os.makedirs("evaluate_result", exist_ok=True)
    res_addr = os.path.join("evaluate_result", prefix)
    np.savetxt(res_addr + 'scales.txt' + tag, res["scales"])
    np.savetxt(res_addr + 'path.txt' + tag, res["poses"])
    return res


if __name__ == "__main__":
    main()
