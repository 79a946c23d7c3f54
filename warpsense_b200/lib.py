"""ctypes binding of libwarpsense_b200.so (the C ABI declared in include/warpsense_b200.h).

The CUDA library is the only compute path: if it cannot be built or loaded this module raises --
there is no CPU fallback anywhere in the package.
"""
import ctypes as C
import os

from . import build as _build

_lib = None

WS_OK = 0
WS_ERR_INVALID = -1
WS_ERR_CUDA = -2
WS_ERR_CAPACITY = -3
WS_ERR_STATE = -4
WS_MAX_POINTS = 1 << 24
WS_REG_DEVICE_SOLVE = 0
WS_TRACK_REFERENCE_POSE = 1
WS_REG_HOST_SOLVE = 1

TIMER_MARCH, TIMER_MERGE, TIMER_REG, TIMER_REPLAY, TIMER_UPDATE = 0, 1, 2, 3, 4


class UpdateCounters(C.Structure):
    _fields_ = [("n_points", C.c_int64), ("n_candidates", C.c_int64), ("n_touched", C.c_int64),
                ("n_written", C.c_int64), ("n_touched_bricks", C.c_int64), ("n_parked", C.c_int64),
                ("n_rounds", C.c_int64), ("n_list", C.c_int64), ("n_record_chunks", C.c_int64),
                ("replay_phase_ns", C.c_int64 * 3)]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_}
        return {k: (int(v) if isinstance(v, int) else [int(x) for x in v]) for k, v in d.items()}


class MapMeta(C.Structure):
    _fields_ = [("tau", C.c_int32), ("map_size", C.c_int32 * 3), ("max_distance", C.c_float),
                ("map_resolution", C.c_int32), ("max_weight", C.c_int32)]


class WarpsenseError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("warpsense_b200 error %d: %s" % (code, msg))
        self.code = code


def _signatures():
    i32p = C.POINTER(C.c_int32)
    i64p = C.POINTER(C.c_int64)
    u32p = C.POINTER(C.c_uint32)
    f32p = C.POINTER(C.c_float)
    vp = C.c_void_p
    hp = C.c_void_p
    return {
        "ws_version": (C.c_char_p, []),
        "ws_create": (C.c_int, [i32p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(hp)]),
        "ws_create_sharded": (C.c_int, [i32p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                        C.c_int32, C.POINTER(hp)]),
        "ws_destroy": (None, [hp]),
        "ws_last_error": (C.c_char_p, [hp]),
        "ws_set_stream": (C.c_int, [hp, vp]),
        "ws_get_stream": (vp, [hp]),
        "ws_sync": (C.c_int, [hp]),
        "ws_map_upload": (C.c_int, [hp, vp, i32p, i32p, i32p]),
        "ws_map_download": (C.c_int, [hp, vp]),
        "ws_map_set_params": (C.c_int, [hp, i32p, i32p]),
        "ws_map_get_params": (C.c_int, [hp, i32p, i32p, i32p]),
        "ws_map_fill": (C.c_int, [hp, C.c_int32, C.c_int32]),
        "ws_map_get_voxel": (C.c_int, [hp, C.c_int32, C.c_int32, C.c_int32, u32p]),
        "ws_map_checksum": (C.c_int, [hp, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_uint64)]),
        "ws_map_set_voxel": (C.c_int, [hp, C.c_int32, C.c_int32, C.c_int32, C.c_uint32]),
        "ws_update_tsdf": (C.c_int, [hp, vp, C.c_int64, i32p, i32p]),
        "ws_update_tsdf_device": (C.c_int, [hp, vp, C.c_int64, i32p, i32p]),
        "ws_get_update_counters": (C.c_int, [hp, C.POINTER(UpdateCounters)]),
        "ws_preprocess_scan": (C.c_int, [hp, vp, C.c_int64, C.c_int32, C.c_int32, f32p, C.c_int32, vp, i64p]),
        "ws_scan_points_device": (vp, [hp, i64p]),
        "ws_voxelgrid_subsample": (C.c_int, [hp, vp, C.c_int64, C.c_int32, C.c_int32, C.c_float, vp, vp, i64p]),
        "ws_update_tsdf_from_ros": (C.c_int, [hp, vp, C.c_int64, C.c_int32, C.c_int32, C.POINTER(C.c_double), f32p, i64p]),
        "ws_reg_prepare": (C.c_int, [hp, vp, C.c_int64]),
        "ws_reg_prepare_device": (C.c_int, [hp, vp, C.c_int64]),
        "ws_reg_step": (C.c_int, [hp, f32p, C.c_int32, i64p, i64p, i32p, i32p]),
        "ws_register_cloud": (C.c_int, [hp, vp, C.c_int64, f32p, C.c_int32, C.c_float, C.c_float,
                                        C.c_int32, C.c_int32, f32p, i32p]),
        "ws_track_scan": (C.c_int, [hp, vp, C.c_int64, C.c_int32, f32p, C.c_int32, C.c_float, C.c_float, C.c_int32,
                                    f32p, f32p, i32p]),
        "ws_track_scan_ex": (C.c_int, [hp, vp, C.c_int64, C.c_int32, f32p, f32p, C.c_int32, C.c_float, C.c_float, C.c_int32,
                                       C.c_int32, f32p, f32p, i32p]),
        "ws_track_submit": (C.c_int, [hp, vp, C.c_int64, C.c_int32, f32p, f32p, C.c_int32, C.c_float, C.c_float, C.c_int32,
                                      C.c_int32, i32p]),
        "ws_track_wait": (C.c_int, [hp, C.c_int32, f32p, f32p, i32p]),
        "ws_reg_get_trace": (C.c_int, [hp, i64p, C.c_int32]),
        "ws_reg_points_device": (vp, [hp, i64p]),
        "ws_reg_begin": (C.c_int, [hp, f32p]),
        "ws_reg_accumulate": (C.c_int, [hp, C.c_int32]),
        "ws_reg_sums_device": (vp, [hp]),
        "ws_reg_sums_get": (C.c_int, [hp, i64p]),
        "ws_reg_sums_set": (C.c_int, [hp, i64p]),
        "ws_peer_export": (C.c_int, [hp, C.POINTER(C.c_uint8)]),
        "ws_peer_attach_ipc": (C.c_int, [hp, C.POINTER(C.c_uint8), C.c_int32]),
        "ws_peer_local_ptr": (vp, [hp]),
        "ws_peer_attach_ptrs": (C.c_int, [hp, C.POINTER(vp), i32p, C.c_int32]),
        "ws_peer_set_timeout": (C.c_int, [hp, C.c_double]),
        "ws_slab_layout": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, i32p, i32p, i32p, C.c_int32]),
        "ws_shard_layout": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_uint8), C.POINTER(C.c_uint8), C.c_int32]),
        "ws_shard_columns": (C.c_int, [hp, C.POINTER(C.c_uint8), C.POINTER(C.c_uint8), C.c_int32]),
        "ws_reg_solve": (C.c_int, [hp, C.c_float, C.c_float]),
        "ws_reg_peek": (C.c_int, [hp, i32p, i32p]),
        "ws_reg_finish": (C.c_int, [hp, f32p, i32p, i32p]),
        "ws_test_reduce": (C.c_int, [hp, i64p, i32p, C.c_int64, i64p, i64p, i32p, i32p]),
        "ws_shift": (C.c_int, [hp, i32p]),
        "ws_write_back": (C.c_int, [hp]),
        "ws_store_num_chunks": (C.c_int64, [hp]),
        "ws_store_chunk_list": (C.c_int, [hp, i32p, C.c_int64]),
        "ws_store_get_chunk": (C.c_int, [hp, C.c_int32, C.c_int32, C.c_int32, u32p]),
        "ws_store_set_chunk": (C.c_int, [hp, C.c_int32, C.c_int32, C.c_int32, u32p]),
        "ws_store_configure": (C.c_int, [hp, C.c_int64]),
        "ws_store_evictions": (C.c_int64, [hp]),
        "ws_export_hdf5": (C.c_int, [hp, C.c_char_p, C.POINTER(MapMeta), f32p, C.c_int64]),
        "ws_hdf5_write_chunks": (C.c_int, [C.c_char_p, C.POINTER(MapMeta), i32p, u32p, C.c_int64, f32p, C.c_int64]),
        "ws_import_hdf5": (C.c_int, [hp, C.c_char_p, C.POINTER(MapMeta), f32p, C.c_int64, i64p, i64p]),
        "ws_map_reload": (C.c_int, [hp]),
        "ws_launch_count": (C.c_int64, [hp]),
        "ws_profile_enable": (C.c_int, [hp, C.c_int32]),
        "ws_profile_reset": (C.c_int, [hp]),
        "ws_profile_get": (C.c_int, [hp, C.c_int32, C.POINTER(C.c_double), i64p]),
        "ws_profile_timeline": (C.c_int64, [hp, C.POINTER(C.c_double), C.c_int64]),
    }


EXPORTED_SYMBOLS = tuple(_signatures().keys())


def load(build_if_needed=True):
    """Load (building first if the sources are newer) libwarpsense_b200.so.  Raises on failure."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("WS_LIB_PATH") or _build.LIB_PATH     # WS_LIB_PATH: A/B runs of alternative builds (tools/)
    if path == _build.LIB_PATH and build_if_needed and _build.needs_build():
        _build.build_library()
    if not os.path.exists(path):
        raise RuntimeError("libwarpsense_b200.so is missing: run `python -m warpsense_b200.build` "
                           "(the CUDA library is the only compute path; there is no fallback)")
    L = C.CDLL(path)
    for name, (res, args) in _signatures().items():
        if os.environ.get("WS_LIB_PATH") and not hasattr(L, name):
            continue               # A/B runs against older builds (tools/gpu_ab.sh) only
        fn = getattr(L, name)      # AttributeError here == ABI/header drift: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L
