import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvoscalerecovery_b200.batch import ScaleRecovery
eng = ScaleRecovery(absolute_reference=1.7)
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 300
pts = np.stack([rng.uniform(0, 1241, n), rng.uniform(186, 376, n)], 1).astype(np.float32)
off = torch.tensor([0, n], dtype=torch.int32, device=eng.device)
out = eng.delaunay_frames(off, torch.from_numpy(pts[:, 0].copy()).to(eng.device), torch.from_numpy(pts[:, 1].copy()).to(eng.device), n)
torch.cuda.synchronize()
print("ntri", out["n_tri"].cpu().numpy(), "status", out["status"].cpu().numpy())
from scipy.spatial import Delaunay
s = np.sort(Delaunay(pts.astype(np.float64)).simplices, 1); s = s[np.lexsort((s[:, 2], s[:, 1], s[:, 0]))]
t = out["tri"].cpu().numpy()[: int(out["n_tri"][0])]
print("equal", np.array_equal(t, s), s.shape)
