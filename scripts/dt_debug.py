"""Debug aid: run the degenerate Delaunay sets one by one and report status / mismatches."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvoscalerecovery_b200.batch import ScaleRecovery
from oracle import exact
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_parity import _degenerate_sets
eng = ScaleRecovery(absolute_reference=1.7)
dev = eng.device
sets = _degenerate_sets()
only = sys.argv[1:] or list(sets)
for k in only:
    p = sets[k]
    off = np.array([0, p.shape[0]], np.int32)
    out = eng.delaunay_frames(torch.from_numpy(off).to(dev), torch.from_numpy(np.ascontiguousarray(p[:, 0])).to(dev),
                              torch.from_numpy(np.ascontiguousarray(p[:, 1])).to(dev), int(p.shape[0]))
    torch.cuda.synchronize()
    nt = int(out["n_tri"].cpu().numpy()[0]); st = int(out["status"].cpu().numpy()[0])
    ref, dup = exact.delaunay_exact(p)
    got = out["tri"].cpu().numpy()[:nt]
    same = got.shape == ref.shape and np.array_equal(got, ref)
    print(k, "n", p.shape[0], "status", st, "ntri", nt, "ref", ref.shape[0], "same", same, flush=True)
