"""The C-ABI library: loads without a GPU, exports every function include/mvosr.h declares, agrees with the Python
struct mirrors, and refuses to compute without a device (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mvosr.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mvosr_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    from mvoscalerecovery_b200 import _native as N
    lib = N.lib()
    names = _declared_functions()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), "libmvosr.so does not export %s" % n
    assert sorted(N.SYMBOLS) == names, "python binding list and header disagree"
    assert lib.mvosr_version() == 200
    assert lib.mvosr_error_string(0) == b"ok" and b"capacity" in lib.mvosr_error_string(-4)


def test_default_config_matches_reference_constants():
    """The constants hard-coded in the reference."""
    from mvoscalerecovery_b200 import _native as N
    c = N.default_config()
    assert c.absolute_reference == 1.75 and c.vanish == 185.0                      # param.py:36, rescale.py:30
    assert (c.fx, c.fy, c.cx, c.cy) == (718.856, 718.856, 607.1928, 185.2157)      # param.py:30-35
    assert (c.min_features, c.min_kept, c.min_selected) == (100, 10, 12)           # param.py:37, rescale.py:133,152
    assert (c.ransac_iterations, c.ransac_threshold, c.ransac_goal_fraction, c.ransac_stop_at_goal) == (100, 0.005, 0.8, 1)
    assert (c.slew_limit, c.window_size, c.height_level_factor) == (0.3, 5, 0.9)
    assert c.triangulation_max_depth == 100.0
    # loose/tight thresholds: largest double s with asin(s)*180/pi < -80 / -85 (rescale.py:85-86 compare degrees)
    for s, deg in ((c.sin_loose, -80.0), (c.sin_tight, -85.0)):
        assert np.degrees(np.arcsin(s)) < deg <= np.degrees(np.arcsin(np.nextafter(s, 0.0)))
    # graph vote table: bit (idx*3+k) <=> p_k(idx) > 0.6 == both edges at vertex k consistent (graph.py:131-145)
    for idx in range(8):
        a, b, cc = (idx >> 2) & 1, (idx >> 1) & 1, idx & 1
        for k, want in enumerate((a and cc, a and b, b and cc)):
            assert ((c.graph_pass_mask >> (idx * 3 + k)) & 1) == int(bool(want))


def test_struct_layouts_match_header():
    from mvoscalerecovery_b200 import _native as N
    from mvoscalerecovery_b200.batch import stats_to_numpy          # asserts itemsize == sizeof(FrameStats) when used
    assert C.sizeof(N.FrameStats) == 14 * 4 + 8 + 32 + 8
    assert C.sizeof(N.DebugBuffers) == 8 * 8
    assert C.sizeof(N.Config) % 8 == 0


def test_no_device_is_an_error_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from mvoscalerecovery_b200 import _native as N
    h = C.c_void_p()
    rc = N.lib().mvosr_create(None, 0, C.byref(h))
    assert rc == -5 and not h.value                                  # MVOSR_E_NO_DEVICE
    from mvoscalerecovery_b200.batch import ScaleRecovery
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ScaleRecovery()


def test_missing_library_raises(monkeypatch):
    from mvoscalerecovery_b200 import _native as N
    monkeypatch.setattr(N, "_lib", None)
    monkeypatch.setattr(N, "LIB_PATH", os.path.join(ROOT, "does_not_exist.so"))
    with pytest.raises(N.NativeLibraryError, match="no CPU fallback"):
        N.lib()


def test_binding_argument_counts_match_the_header():
    """Every prototype of include/mvosr.h against the ctypes argtypes of the binding: same number of parameters, and pointers /
    integers / doubles in the same places (a mismatch here is a crash on the GPU box, not an exception)."""
    from mvoscalerecovery_b200 import _native as N
    lib = N.lib()
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    checked = 0
    for m in re.finditer(r"\b(mvosr_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        name, params = m.group(1), m.group(2).strip()
        fn = getattr(lib, name)
        if fn.argtypes is None:
            continue
        plist = [] if params in ("", "void") else [p.strip() for p in params.split(",")]
        assert len(plist) == len(fn.argtypes), (name, len(plist), len(fn.argtypes))
        for p, t in zip(plist, fn.argtypes):
            is_ptr = "*" in p
            if is_ptr:
                assert t in (C.c_void_p, C.c_char_p) or hasattr(t, "contents") or getattr(t, "_type_", None) is not None, (name, p, t)
            elif p.startswith("double"):
                assert t is C.c_double, (name, p, t)
            elif p.startswith("uint64_t"):
                assert t is C.c_uint64, (name, p, t)
            elif p.startswith(("int32_t", "int ")):
                assert t in (C.c_int32, C.c_int), (name, p, t)
        checked += 1
    assert checked >= 18
