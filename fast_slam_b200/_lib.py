"""ctypes binding of libfs2.so (include/fs2.h).  There is NO fallback: if the CUDA library is missing or
does not load, importing a filter fails loudly with instructions to build it."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FS2_LIB") or os.path.join(_HERE, "libfs2.so")   # FS2_LIB: build-variant experiments

FS2_OK = 0
FS2_STATS_LEN = 16
STAT_TOTAL, STAT_SUMSQ, STAT_NEFF, STAT_WMAX, STAT_ARGMAX, STAT_EST_X, STAT_EST_Y, STAT_EST_YAW = range(8)
STAT_RESAMPLED, STAT_COPIES, STAT_ANOMALY, STAT_STUCK, STAT_ARGMAX_ID, STAT_DEFERRED = 8, 9, 10, 11, 12, 13
ST_SINGULAR_LM, ST_SINGULAR_Q, ST_PDF_FAILED, ST_MAP_FULL = 1, 2, 4, 8
FLAG_FORCE_SEQUENTIAL = 1


class Fs2Config(C.Structure):
    _fields_ = [
        ("num_particles", C.c_int64), ("global_particles", C.c_int64), ("global_offset", C.c_int64),
        ("landmark_capacity", C.c_int32), ("device", C.c_int32), ("flags", C.c_int32), ("spare_slots", C.c_int32),
        ("translation_noise", C.c_double), ("rotation_noise", C.c_double),
        ("measurement_noise", C.c_double * 4), ("max_landmark_distance", C.c_double), ("seed", C.c_uint64),
    ]


class Fs2Ptrs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("y", C.c_void_p), ("yaw", C.c_void_p), ("w", C.c_void_p), ("count", C.c_void_p),
        ("lm", C.c_void_p), ("status", C.c_void_p), ("noise", C.c_void_p), ("cumsum", C.c_void_p),
        ("ancestor", C.c_void_p), ("stats", C.c_void_p), ("num_particles", C.c_int64),
        ("landmark_capacity", C.c_int32), ("reserved", C.c_int32),
    ]


class Fs2StepResult(C.Structure):
    _fields_ = [("x", C.c_double), ("y", C.c_double), ("yaw", C.c_double), ("neff", C.c_double),
                ("total", C.c_double), ("resampled", C.c_int32), ("status_or", C.c_int32)]


class Fs2KlInfo(C.Structure):
    _fields_ = [("n_points", C.c_int64), ("min_samples", C.c_int64), ("involved_points", C.c_int64),
                ("noise_points", C.c_int64), ("tiles", C.c_int32), ("clusters", C.c_int32), ("err_bits", C.c_int32),
                ("skipped", C.c_int32), ("tiles_used", C.c_int32), ("reserved", C.c_int32)]


class Fs2Error(RuntimeError):
    pass


_lib = None

# every symbol include/fs2.h declares (tests check the library exports exactly these)
EXPORTS = [
    "fs2_abi_version", "fs2_strerror", "fs2_last_cuda_error", "fs2_create", "fs2_destroy", "fs2_reset",
    "fs2_get_ptrs", "fs2_draw_noise", "fs2_motion", "fs2_update", "fs2_motion_update", "fs2_weight_total",
    "fs2_normalize", "fs2_estimate", "fs2_resample_indices", "fs2_gather", "fs2_gather_ext", "fs2_pack_records", "fs2_ipc_export", "fs2_ipc_open_peers", "fs2_gather_p2p", "fs2_gather_commit", "fs2_place_enable", "fs2_place_logical_ids", "fs2_place_resample", "fs2_place_commit", "fs2_pull_records", "fs2_finish_step", "fs2_sync_maps", "fs2_step_host", "fs2_launch_count",
    "fs2_upload_state", "fs2_download_state", "fs2_download_particles", "fs2_debug_obs_batch_size", "fs2_debug_obs_batch", "fs2_frontend", "fs2_frontend_max_measurements",
    "fs2_known_landmarks", "fs2_cluster_points", "fs2_frontend_polar", "fs2_frontend_release", "fs2_line_filter", "fs2_icp", "fs2_hough_intersections", "fs2_hough_max_intersections", "fs2_frontend_sq_threshold",
    "fs2_kl_record_bytes", "fs2_kl_shard_begin", "fs2_kl_shard_set_bases", "fs2_kl_shard_count", "fs2_kl_shard_export", "fs2_kl_shard_merge",
    "fs2_kl_shard_extract", "fs2_kl_shard_finish",
]


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Fs2Error(
            "fast_slam_b200: %s not found. The filter has no CPU fallback; build the CUDA library first "
            "(python -c 'import __graft_entry__ as g; g.build()' or make -C fast_slam_b200/csrc)." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, u64, d = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double
    pd = C.POINTER(C.c_double)
    L.fs2_abi_version.restype = C.c_int
    L.fs2_strerror.restype = C.c_char_p
    L.fs2_strerror.argtypes = [C.c_int]
    L.fs2_last_cuda_error.restype = C.c_char_p
    L.fs2_create.argtypes = [C.POINTER(Fs2Config), C.POINTER(vp)]
    L.fs2_destroy.argtypes = [vp]
    L.fs2_reset.argtypes = [vp, vp]
    L.fs2_get_ptrs.argtypes = [vp, C.POINTER(Fs2Ptrs)]
    L.fs2_draw_noise.argtypes = [vp, d, u64, vp, vp]
    L.fs2_motion.argtypes = [vp, d, d, vp, vp]
    L.fs2_update.argtypes = [vp, pd, i32, vp, vp]
    L.fs2_motion_update.argtypes = [vp, d, d, vp, pd, i32, vp, vp]
    L.fs2_weight_total.argtypes = [vp, vp]
    L.fs2_normalize.argtypes = [vp, vp, vp]
    L.fs2_estimate.argtypes = [vp, vp]
    L.fs2_resample_indices.argtypes = [vp, vp, i64, d, i64, i64, vp, vp]
    L.fs2_gather.argtypes = [vp, vp, vp]
    L.fs2_gather_ext.argtypes = [vp, vp, vp, i64, vp]
    L.fs2_pack_records.argtypes = [vp, vp, i64, vp, vp]
    L.fs2_ipc_export.argtypes = [vp, vp]
    L.fs2_ipc_open_peers.argtypes = [vp, vp, i32, i32]
    L.fs2_pull_records.argtypes = [vp, i32, vp, i64, vp, vp]
    L.fs2_gather_p2p.argtypes = [vp, vp, vp]
    L.fs2_gather_commit.argtypes = [vp, vp]
    L.fs2_place_enable.argtypes = [vp, vp]
    L.fs2_place_logical_ids.restype = vp
    L.fs2_place_logical_ids.argtypes = [vp]
    L.fs2_place_resample.argtypes = [vp, vp, vp, vp, C.POINTER(C.c_int64), vp]
    L.fs2_place_commit.argtypes = [vp, vp]
    L.fs2_finish_step.argtypes = [vp, d, vp, vp]
    L.fs2_sync_maps.argtypes = [vp, vp]
    L.fs2_step_host.argtypes = [vp, d, d, pd, i32, pd, u64, d, vp, vp, C.POINTER(Fs2StepResult), vp]
    L.fs2_launch_count.restype = i64
    L.fs2_launch_count.argtypes = [vp]
    L.fs2_upload_state.argtypes = [vp, pd, pd, pd, pd, C.POINTER(C.c_int32), pd, vp]
    L.fs2_download_state.argtypes = [vp, pd, pd, pd, pd, C.POINTER(C.c_int32), pd, C.POINTER(C.c_int32), vp]
    L.fs2_download_particles.argtypes = [vp, C.POINTER(C.c_int64), i64, pd, pd, pd, pd, C.POINTER(C.c_int32), pd, vp]
    L.fs2_debug_obs_batch.argtypes = [pd, i32, vp]
    L.fs2_frontend.argtypes = [pd, i32, i32, d, i32, pd, C.POINTER(C.c_int32), C.POINTER(C.c_int32), vp]
    L.fs2_frontend_release.argtypes = [i32]
    L.fs2_hough_intersections.argtypes = [pd, i32, i32, i32, C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_int32), vp]
    L.fs2_icp.argtypes = [pd, pd, i32, i32, i32, i32, d, i32, pd, pd, C.POINTER(C.c_int32), vp]
    L.fs2_line_filter.argtypes = [pd, i32, i32, d, i32, pd, vp]
    L.fs2_frontend_polar.argtypes = [pd, pd, i32, i32, d, d, d, i32, pd, C.POINTER(C.c_int32), C.POINTER(C.c_int32), vp]
    pi64 = C.POINTER(C.c_int64)
    L.fs2_kl_shard_begin.argtypes = [vp, d, pi64, vp]
    L.fs2_kl_shard_set_bases.argtypes = [vp, vp, vp]
    L.fs2_kl_shard_count.argtypes = [vp, i64, C.POINTER(C.c_int32), vp]
    L.fs2_kl_shard_export.argtypes = [vp, vp, i32, vp]
    L.fs2_kl_shard_merge.argtypes = [vp, vp, i32, i64, pi64, vp]
    L.fs2_kl_shard_extract.argtypes = [vp, vp, i64, pi64, vp]
    L.fs2_kl_shard_finish.argtypes = [vp, vp, i64, i64, i32, pd, pi64, C.POINTER(C.c_int32), C.POINTER(Fs2KlInfo), vp]
    L.fs2_known_landmarks.argtypes = [vp, d, d, i64, i32, pd, C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(Fs2KlInfo), vp]
    L.fs2_cluster_points.argtypes = [pd, i64, d, i64, i32, i32, pd, C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(Fs2KlInfo)]
    for name in EXPORTS:
        getattr(L, name)
    _lib = L
    return L


def check(status: int, what: str = "fs2 call"):
    if status != FS2_OK:
        L = load()
        msg = L.fs2_strerror(status).decode()
        cuda = L.fs2_last_cuda_error().decode()
        raise Fs2Error("%s failed: %s%s" % (what, msg, (" [" + cuda + "]") if cuda else ""))
