import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


class Golden:
    """One golden sequence written by tests/golden/make_golden.py (outputs of the reference itself)."""

    def __init__(self, name):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.n_frames = int(self.z["n_frames"])
        self.seed = int(self.z["seed"])

    def f3(self, f):
        return self.z["f%d_f3" % f]

    def f2(self, f):
        return self.z["f%d_f2" % f]

    def called(self, f):
        return bool(self.z["f%d_called" % f])

    def get(self, f, key):
        return self.z["f%d_%s" % (f, key)]

    def scalars(self, f):
        s = self.z["f%d_scalars" % f]
        return dict(height_level=s[0], best_ic=int(s[1]), raw_scale=s[2], height=s[3], updated=bool(s[4]),
                    state_before=s[5], state_after=s[6], scale_out=s[7], second_dt=bool(s[8]))


GOLDEN_NAMES = ["seq_2k", "seq_small", "seq_clean"]


@pytest.fixture(scope="session", params=GOLDEN_NAMES)
def golden(request):
    return Golden(request.param)


@pytest.fixture(scope="session")
def engine():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from mvoscalerecovery_b200.batch import ScaleRecovery
    return ScaleRecovery(absolute_reference=1.7)
