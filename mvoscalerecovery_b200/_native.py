"""ctypes binding of libmvosr.so (include/mvosr.h).  There is no fallback: if the CUDA library
is missing or cannot be loaded every entry point of this package raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libmvosr.so")

OK = 0
ST_UPDATED, ST_SECOND_DT, ST_FEW_ROI, ST_NO_MODEL, ST_BAD_INPUT, ST_OVERFLOW, ST_SKIPPED, ST_SINGULAR = 1, 2, 4, 8, 16, 32, 64, 128


class Config(C.Structure):
    _fields_ = [
        ("absolute_reference", C.c_double),
        ("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
        ("vanish", C.c_float),
        ("min_features", C.c_int32), ("min_kept", C.c_int32), ("min_selected", C.c_int32),
        ("sin_loose", C.c_double), ("sin_tight", C.c_double), ("height_level_factor", C.c_double),
        ("ransac_iterations", C.c_int32), ("ransac_stop_at_goal", C.c_int32),
        ("ransac_threshold", C.c_double), ("ransac_goal_fraction", C.c_double),
        ("graph_pass_mask", C.c_uint32),
        ("slew_limit", C.c_double), ("window_size", C.c_int32),
        ("triangulation_max_depth", C.c_double),
        ("reserved", C.c_int32 * 8),
    ]


class FrameStats(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "n_features", "n_roi", "n_dup", "n_kept", "n_tri", "n_loose", "n_tight", "n_valid", "best_hyp", "best_ic",
        "hyps_used", "n_degenerate", "n_deferred", "n_exact")] + [
        ("height_level", C.c_double), ("model", C.c_double * 4), ("height", C.c_double)]


class FrameRecord(C.Structure):
    """mvosr_frame_record: one frame's (raw_scale, n_features, status) in 16 bytes -- a rank's contribution to the fleet gather."""
    _fields_ = [("raw_scale", C.c_double), ("n_features", C.c_int32), ("status", C.c_uint8), ("pad", C.c_uint8 * 3)]


RECORD_BYTES = C.sizeof(FrameRecord)
assert RECORD_BYTES == 16


class DebugBuffers(C.Structure):
    _fields_ = [("tri1", C.c_void_p), ("n_tri1", C.c_void_p), ("keep", C.c_void_p), ("tri2", C.c_void_p),
                ("tri_flags", C.c_void_p), ("tri_height", C.c_void_p), ("inlier", C.c_void_p), ("data_id", C.c_void_p)]


# every symbol include/mvosr.h declares
SYMBOLS = [
    "mvosr_version", "mvosr_error_string", "mvosr_last_cuda_error", "mvosr_default_config", "mvosr_create",
    "mvosr_destroy", "mvosr_get_config", "mvosr_triangulate_frames", "mvosr_scale_frames",
    "mvosr_scale_frames_from_correspondences", "mvosr_filter_sequences", "mvosr_delaunay_frames",
    "mvosr_recover_scales_host", "mvosr_recover_fleet_host", "mvosr_launch_count", "mvosr_set_phase_timing",
    "mvosr_triangle_planes", "mvosr_triangle_votes", "mvosr_ransac_planes", "mvosr_integrate_paths", "mvosr_depth_from_mesh", "mvosr_recover_pose_frames",
    "mvosr_find_essential_frames", "mvosr_pose_mask_frames", "mvosr_bucket_frames",
    "mvosr_scale_shard_from_correspondences", "mvosr_filter_records", "mvosr_scale_frames_f64", "mvosr_scale_frame_host_f64",
]

_lib = None


class NativeLibraryError(RuntimeError):
    pass


def lib():
    """Load libmvosr.so; raises NativeLibraryError (never falls back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise NativeLibraryError(
            "libmvosr.so not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `python -m mvoscalerecovery_b200.build`. There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32, u64 = C.c_void_p, C.c_int32, C.c_uint64
    L.mvosr_version.restype = C.c_int
    L.mvosr_error_string.restype = C.c_char_p
    L.mvosr_error_string.argtypes = [C.c_int]
    L.mvosr_last_cuda_error.restype = C.c_char_p
    L.mvosr_default_config.argtypes = [C.POINTER(Config)]
    L.mvosr_create.argtypes = [C.POINTER(Config), C.c_int, C.POINTER(vp)]
    L.mvosr_destroy.argtypes = [vp]
    L.mvosr_get_config.argtypes = [vp, C.POINTER(Config)]
    L.mvosr_triangulate_frames.argtypes = [vp, i32] + [vp] * 13 + [vp]
    L.mvosr_scale_frames.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, u64, vp, vp, vp,
                                     C.POINTER(DebugBuffers), vp]
    L.mvosr_scale_frames_from_correspondences.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, u64,
                                                          vp, vp, vp, vp, vp]
    L.mvosr_filter_sequences.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, vp, vp]
    L.mvosr_scale_shard_from_correspondences.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, vp, i32, vp, vp, i32, i32, vp, u64, vp, vp]
    L.mvosr_filter_records.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, vp]
    L.mvosr_scale_frames_f64.argtypes = [vp, i32, vp, vp, vp, i32, i32, i32, u64, vp, vp, vp, C.POINTER(DebugBuffers), vp]
    L.mvosr_scale_frame_host_f64.argtypes = [vp, i32, vp, vp, i32, i32, u64, C.POINTER(FrameRecord), C.POINTER(FrameStats)]
    L.mvosr_delaunay_frames.argtypes = [vp, i32, vp, vp, vp, i32, vp, vp, vp, vp]
    L.mvosr_recover_scales_host.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, vp, i32, i32, u64, vp, vp, vp]
    L.mvosr_recover_fleet_host.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, u64, vp, vp, vp]
    L.mvosr_launch_count.argtypes = [vp]
    f64 = C.c_double
    L.mvosr_triangle_planes.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp]
    L.mvosr_triangle_votes.argtypes = [vp, i32, vp, vp, vp, i32, vp, vp, vp]
    L.mvosr_ransac_planes.argtypes = [vp, i32, vp, vp, i32, f64, f64, i32, u64, vp, i32, vp, vp, vp, vp, vp]
    L.mvosr_integrate_paths.argtypes = [vp, i32, vp, vp, vp, vp, vp]
    L.mvosr_recover_pose_frames.argtypes = [vp, i32] + [vp] * 10
    L.mvosr_bucket_frames.argtypes = [vp, i32, vp, vp, vp, i32, i32, u64, vp, i32, vp, vp, vp, vp]
    L.mvosr_pose_mask_frames.argtypes = [vp, i32] + [vp] * 9
    L.mvosr_find_essential_frames.argtypes = [vp, i32] + [vp] * 5 + [i32, f64, f64, u64, vp, i32] + [vp] * 6
    L.mvosr_depth_from_mesh.argtypes = [vp, i32, i32, f64, f64, f64, f64, i32, vp, vp, vp, vp, vp, vp]
    L.mvosr_set_phase_timing.argtypes = [vp, vp]
    L.mvosr_launch_count.restype = C.c_int64
    for s in SYMBOLS:
        getattr(L, s)
    _lib = L
    return L


def check(rc: int):
    if rc != OK:
        L = lib()
        msg = L.mvosr_error_string(rc).decode()
        if rc == -2:
            msg += ": " + L.mvosr_last_cuda_error().decode()
        raise RuntimeError("libmvosr: %s (%d)" % (msg, rc))


def default_config() -> Config:
    c = Config()
    check(lib().mvosr_default_config(C.byref(c)))
    return c
