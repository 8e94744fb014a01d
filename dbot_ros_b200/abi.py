"""ctypes binding of libdbot_b200.so -- one Python function per entry point of include/dbot_b200.h.

This is the stub a maintainer of the reference would write for its FFI (see INTEGRATION.md for the
C++ one); nothing here computes anything.  A missing library is a hard error (no fallback).
"""
import ctypes as C
import os

import numpy as np

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libdbot_b200.so")
MAX_OBJECTS = 8
KERNEL_NAMES = ("observation", "loglikes", "update", "render", "plan")


class DbotB200Error(RuntimeError):
    def __init__(self, code, message):
        super().__init__("dbot_b200 error %d: %s" % (code, message))
        self.code = code


class Config(C.Structure):
    """dbot_b200_config (include/dbot_b200.h)"""

    _fields_ = [
        ("rows", C.c_int32),
        ("cols", C.c_int32),
        ("K", C.c_double * 9),
        ("max_particles", C.c_int32),
        ("n_objects", C.c_int32),
        ("vertices", C.POINTER(C.POINTER(C.c_float))),
        ("n_vertices", C.POINTER(C.c_int32)),
        ("triangles", C.POINTER(C.POINTER(C.c_int32))),
        ("n_triangles", C.POINTER(C.c_int32)),
        ("tail_weight", C.c_double),
        ("model_sigma", C.c_double),
        ("sigma_factor", C.c_double),
        ("half_life_depth", C.c_double),
        ("max_depth", C.c_double),
        ("p_occluded_visible", C.c_double),
        ("p_occluded_occluded", C.c_double),
        ("initial_occlusion_prob", C.c_double),
        ("delta_time", C.c_double),
        ("device_id", C.c_int32),
        ("occlusion_mode", C.c_int32),
        ("cull_backfaces", C.c_int32),
        ("n_devices", C.c_int32),
        ("device_ids", C.POINTER(C.c_int32)),
        ("gather_mode", C.c_int32),
    ]


class FilterConfig(C.Structure):
    """dbot_b200_filter_config (include/dbot_b200.h)"""

    _fields_ = [
        ("linear_sigma", C.c_double * 3),
        ("angular_sigma", C.c_double * 3),
        ("velocity_factor", C.c_double),
        ("max_kl_divergence", C.c_double),
        ("seed", C.c_uint64),
        ("n_particles", C.c_int32),
        ("moving_average_update_rate", C.c_double),
        ("centers", (C.c_double * 3) * MAX_OBJECTS),
    ]


IPC_HANDLE_BYTES = 64   # DBOT_B200_IPC_HANDLE_BYTES
OPT_STAGING_THREADS = 1  # DBOT_B200_OPT_STAGING_THREADS

# name -> (restype, argtypes); every symbol include/dbot_b200.h declares
_P = C.c_void_p
SIGNATURES = {
    "dbot_b200_create": (C.c_int, [C.POINTER(Config), C.POINTER(_P)]),
    "dbot_b200_destroy": (C.c_int, [_P]),
    "dbot_b200_reset": (C.c_int, [_P]),
    "dbot_b200_set_observation": (C.c_int, [_P, _P]),
    "dbot_b200_loglikes": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, _P]),
    "dbot_b200_render": (C.c_int, [_P, _P, _P]),
    "dbot_b200_set_observation_device": (C.c_int, [_P, _P]),
    "dbot_b200_loglikes_device": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, _P]),
    "dbot_b200_stream": (_P, [_P]),
    "dbot_b200_synchronize": (C.c_int, [_P]),
    "dbot_b200_slot_floats": (C.c_int64, [_P]),
    "dbot_b200_export_slot_device": (C.c_int, [_P, C.c_int32, _P]),
    "dbot_b200_import_slot_device": (C.c_int, [_P, _P, C.POINTER(C.c_int32)]),
    "dbot_b200_filter_create": (C.c_int, [_P, C.POINTER(FilterConfig), C.POINTER(_P)]),
    "dbot_b200_filter_destroy": (C.c_int, [_P]),
    "dbot_b200_filter_initialize": (C.c_int, [_P, _P]),
    "dbot_b200_filter_track": (C.c_int, [_P, _P, _P, C.POINTER(C.c_int32)]),
    "dbot_b200_update_empty": (C.c_int, [_P]),
    "dbot_b200_peer_gather_export": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, _P]),
    "dbot_b200_peer_gather_connect": (C.c_int, [_P, _P]),
    "dbot_b200_loglikes_gather": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, C.c_int32, _P]),
    "dbot_b200_loglikes_gather_device": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, _P, C.c_int32]),
    "dbot_b200_peer_gather_result": (_P, [_P]),
    "dbot_b200_shard_scan": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32)]),
    "dbot_b200_set_option": (C.c_int, [_P, C.c_int32, C.c_int64]),
    "dbot_b200_device_count": (C.c_int, [_P]),
    "dbot_b200_migrated_slots": (C.c_int64, [_P]),
    "dbot_b200_n_slots": (C.c_int, [_P]),
    "dbot_b200_cull_sign": (C.c_int, [_P, C.c_int32]),
    "dbot_b200_observation_time": (C.c_double, [_P]),
    "dbot_b200_slot_rows": (C.c_int, [_P, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "dbot_b200_slot_rect": (C.c_int, [_P, C.c_int32, C.POINTER(C.c_int32)]),
    "dbot_b200_free_slot_count": (C.c_int, [_P]),
    "dbot_b200_get_occlusion": (C.c_int, [_P, C.c_int32, _P]),
    "dbot_b200_timer_begin": (C.c_int, [_P]),
    "dbot_b200_timer_end": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "dbot_b200_profile_enable": (C.c_int, [_P, C.c_int32]),
    "dbot_b200_profile_read": (C.c_int, [_P, _P, _P]),
    "dbot_b200_launch_count": (C.c_int64, [_P]),
    "dbot_b200_device_bytes": (C.c_int64, [_P]),
    "dbot_b200_abi_version": (C.c_int, []),
    "dbot_b200_last_error": (C.c_char_p, []),
}

_lib = None


def load_library(path=None):
    """dlopen libdbot_b200.so and type every entry point.  Raises if the library is not built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("DBOT_B200_LIB") or LIB_PATH  # env override: profiling variants
    if not os.path.exists(p):
        raise ImportError(
            "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(dbot_ros_b200 has no CPU fallback)" % p
        )
    lib = C.CDLL(p)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def check(lib, rc):
    if rc != 0:
        raise DbotB200Error(rc, lib.dbot_b200_last_error().decode("utf-8", "replace"))


def make_config(rows, cols, K, max_particles, vertices, triangles, *, tail_weight, model_sigma,
                sigma_factor, half_life_depth, max_depth, p_occluded_visible, p_occluded_occluded,
                initial_occlusion_prob, delta_time, device_id=0, occlusion_mode=0,
                cull_backfaces=0, device_ids=None, gather_mode=0, struct=Config):
    """Fill a dbot_b200_config (or a layout-compatible struct).  Returns (cfg, keepalive)."""
    n_obj = len(vertices)
    v_arr = [np.ascontiguousarray(v, dtype=np.float32).reshape(-1, 3) for v in vertices]
    t_arr = [np.ascontiguousarray(t, dtype=np.int32).reshape(-1, 3) for t in triangles]
    cfg = struct()
    cfg.rows, cfg.cols = int(rows), int(cols)
    Kf = np.asarray(K, dtype=np.float64).reshape(9)
    for i in range(9):
        cfg.K[i] = float(Kf[i])
    if hasattr(cfg, "max_particles"):
        cfg.max_particles = int(max_particles)
    cfg.n_objects = n_obj
    vp = (C.POINTER(C.c_float) * n_obj)(*[a.ctypes.data_as(C.POINTER(C.c_float)) for a in v_arr])
    tp = (C.POINTER(C.c_int32) * n_obj)(*[a.ctypes.data_as(C.POINTER(C.c_int32)) for a in t_arr])
    nv = (C.c_int32 * n_obj)(*[a.shape[0] for a in v_arr])
    nt = (C.c_int32 * n_obj)(*[a.shape[0] for a in t_arr])
    cfg.vertices = C.cast(vp, C.POINTER(C.POINTER(C.c_float)))
    cfg.triangles = C.cast(tp, C.POINTER(C.POINTER(C.c_int32)))
    cfg.n_vertices = C.cast(nv, C.POINTER(C.c_int32))
    cfg.n_triangles = C.cast(nt, C.POINTER(C.c_int32))
    cfg.tail_weight = tail_weight
    cfg.model_sigma = model_sigma
    cfg.sigma_factor = sigma_factor
    cfg.half_life_depth = half_life_depth
    cfg.max_depth = max_depth
    cfg.p_occluded_visible = p_occluded_visible
    cfg.p_occluded_occluded = p_occluded_occluded
    if hasattr(cfg, "initial_occlusion_prob"):
        cfg.initial_occlusion_prob = initial_occlusion_prob
    if hasattr(cfg, "delta_time"):
        cfg.delta_time = delta_time
    if hasattr(cfg, "device_id"):
        cfg.device_id = int(device_id)
    if hasattr(cfg, "occlusion_mode"):
        cfg.occlusion_mode = int(occlusion_mode)
    ids = None
    if struct is Config:
        cfg.cull_backfaces = int(cull_backfaces)
        cfg.gather_mode = int(gather_mode)
        if device_ids is not None:
            ids = (C.c_int32 * len(device_ids))(*[int(d) for d in device_ids])
            cfg.n_devices = len(device_ids)
            cfg.device_ids = C.cast(ids, C.POINTER(C.c_int32))
    return cfg, (v_arr, t_arr, vp, tp, nv, nt, ids)


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)
