"""ctypes binding of the C ABI declared in include/pyphy_b200.h.

There is no CPU fallback: if the CUDA library has not been built, importing this
module raises, and so does every product entry point that needs it.
"""
from __future__ import annotations

import ctypes
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "_lib", "libpyphy_b200.so")

PPB_F32, PPB_F64 = 0, 1
PPB_MAX_BALLS, PPB_MAX_WALLS = 32, 16

STATUS_NAMES = {
    0: "OK",
    1: "ASSERT_TCOL_NEG (primitives.py:610)",
    2: "ASSERT_OVERLAP (geometry.py:413)",
    3: "ASSERT_THETA (geometry.py:436-438)",
    4: "ASSERT_THETA_A (geometry.py:447)",
    5: "ASSERT_INTPOINT (dynamics.py:34)",
    6: "TRAP_AMISS (dynamics.py:43-45)",
    7: "MATH_DOMAIN (geometry.py:174,444)",
    8: "ASSERT_TCOL_EQ (primitives.py:664)",
    9: "ITER_LIMIT (engine guard)",
}


class PpbError(RuntimeError):
    pass


class Config(ctypes.Structure):
    _fields_ = [
        ("n_worlds", ctypes.c_int32), ("n_balls", ctypes.c_int32), ("n_walls", ctypes.c_int32),
        ("dtype", ctypes.c_int32), ("canvas_w", ctypes.c_int32), ("canvas_h", ctypes.c_int32),
        ("device", ctypes.c_int32), ("max_substeps", ctypes.c_int32),
        ("dt", ctypes.c_double), ("gx", ctypes.c_double), ("gy", ctypes.c_double),
        ("friction", ctypes.c_double), ("event_capacity", ctypes.c_int64),
        ("order", ctypes.POINTER(ctypes.c_int32)),
    ]


class SamplerParams(ctypes.Structure):
    _fields_ = [("arena", ctypes.c_double), ("w_thick", ctypes.c_double), ("mn_wlen", ctypes.c_double),
                ("mx_wlen", ctypes.c_double), ("mn_ball", ctypes.c_double), ("mx_ball", ctypes.c_double),
                ("mn_force", ctypes.c_double), ("mx_force", ctypes.c_double), ("force_t", ctypes.c_double),
                ("seed", ctypes.c_uint64), ("world_index0", ctypes.c_int64), ("max_tries", ctypes.c_int32),
                ("n_theta", ctypes.c_int32), ("theta", ctypes.c_double * 4)]


class Event(ctypes.Structure):
    _fields_ = [("world", ctypes.c_int32), ("seq", ctypes.c_int32), ("step", ctypes.c_int32),
                ("ball", ctypes.c_int16), ("partner", ctypes.c_int16)]


class BallStyle(ctypes.Structure):
    _fields_ = [("s_thick", ctypes.c_int32), ("fill", ctypes.c_float * 4), ("stroke", ctypes.c_float * 4)]


class WallStyle(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("fill", ctypes.c_float * 4)]


_vp, _i32, _i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int32)

# name -> (restype, argtypes); every entry is a symbol include/pyphy_b200.h declares
SIGNATURES = {
    "ppb_last_error": (ctypes.c_char_p, []),
    "ppb_version": (ctypes.c_int, []),
    "ppb_create": (ctypes.c_int, [ctypes.POINTER(Config), ctypes.POINTER(_vp)]),
    "ppb_destroy": (ctypes.c_int, [_vp]),
    "ppb_set_walls": (ctypes.c_int, [_vp, _i32, _i32, _dp, _vp]),
    "ppb_set_balls": (ctypes.c_int, [_vp, _i32, _i32, _dp, _dp, _dp, _dp, _vp]),
    "ppb_apply_force": (ctypes.c_int, [_vp, _i32, _i32, _dp, ctypes.c_double, _vp]),
    "ppb_sample_rect_worlds": (ctypes.c_int, [_vp, _i32, _i32, ctypes.POINTER(SamplerParams), _ip, _vp]),
    "ppb_update_balls": (ctypes.c_int, [_vp, _i32, _i32, _dp, _dp, _vp]),
    "ppb_requeue": (ctypes.c_int, [_vp, _i32, _i32, _vp]),
    "ppb_get_balls": (ctypes.c_int, [_vp, _i32, _i32, _dp, _dp, _vp]),
    "ppb_get_radius_mass": (ctypes.c_int, [_vp, _i32, _i32, _dp, _dp, _vp]),
    "ppb_get_walls": (ctypes.c_int, [_vp, _i32, _i32, _dp, _vp]),
    "ppb_get_collision_cache": (ctypes.c_int, [_vp, _i32, _i32, _dp, _ip, _vp]),
    "ppb_get_status": (ctypes.c_int, [_vp, _i32, _i32, _ip, _ip, _vp]),
    "ppb_set_styles": (ctypes.c_int, [_vp, ctypes.POINTER(BallStyle), ctypes.POINTER(WallStyle), _i32, _i32, _vp]),
    "ppb_read_sprite": (ctypes.c_int, [_vp, _i32, _i32, _vp, _i64, _ip, _vp]),
    "ppb_set_gravity": (ctypes.c_int, [_vp, ctypes.c_double, ctypes.c_double]),
    "ppb_step_async": (ctypes.c_int, [_vp, _i32, _vp, _vp, _i32, _vp]),
    "ppb_scan_async": (ctypes.c_int, [_vp, _vp]),
    "ppb_get_collision_response": (ctypes.c_int, [_vp, _i32, _i32, _dp, _dp, _vp]),
    "ppb_step_ring_async": (ctypes.c_int, [_vp, _i32, _vp, _vp, _i32, _i32, _vp]),
    "ppb_render_async": (ctypes.c_int, [_vp, _vp, _vp]),
    "ppb_simulate_host": (ctypes.c_int, [_vp, _i32, _vp, _vp, _i32, _vp]),
    "ppb_simulate_host_async": (ctypes.c_int, [_vp, _i32, _vp, _vp, _i32, _vp]),
    "ppb_download_frames": (ctypes.c_int, [_vp, _vp, ctypes.c_int64, _vp, _vp]),
    "ppb_host_sync": (ctypes.c_int, [_vp, _i32, _vp]),
    "ppb_crop_async": (ctypes.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp]),
    "ppb_resize_crops_async": (ctypes.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "ppb_horizon_async": (ctypes.c_int, [_vp, _vp, _i32, _i32, _vp, _vp]),
    "ppb_collision_flags_async": (ctypes.c_int, [_vp, _i64, _i32, ctypes.c_float, _vp, _vp]),
    "ppb_read_events": (ctypes.c_int, [_vp, ctypes.POINTER(Event), _i64, ctypes.POINTER(_i64), _vp]),
    "ppb_clear_events": (ctypes.c_int, [_vp, _vp]),
    "ppb_get_counters": (ctypes.c_int, [_vp, ctypes.POINTER(_i64), _vp]),
    "ppb_device_ptr": (ctypes.c_int, [_vp, _i32, ctypes.POINTER(_vp), ctypes.POINTER(_i64)]),
    "ppb_set_profiling": (ctypes.c_int, [_vp, _i32]),
    "ppb_get_kernel_times": (ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_float), _ip]),
}

_lib = None


def load():
    """dlopen the in-tree CUDA library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "pyphy_engine_b200: %s is missing -- build it with `python -m pyphy_engine_b200.build` "
            "(nvcc, sm_100a).  There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        msg = load().ppb_last_error()
        raise PpbError("ppb error %d: %s" % (rc, msg.decode() if msg else "?"))
