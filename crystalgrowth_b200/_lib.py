"""ctypes loader for libkobayashi_cuda.so — the only compute path of this package.

There is deliberately NO fallback: if the shared library is missing or cannot be loaded, importing the
product API raises.  (The CPU oracle under oracle/ is test infrastructure and is never imported here.)
"""
from __future__ import annotations

import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
# KOB_LIB_PATH: development knob to load an alternative build of the same library (kernel tuning variants)
LIB_PATH = os.environ.get("KOB_LIB_PATH") or os.path.join(PKG, "libkobayashi_cuda.so")

KOB_F32, KOB_F64 = 0, 1
KOB_KERNEL_STRICT, KOB_KERNEL_FAST = 0, 1


class KobParams(C.Structure):
    """kob_params, include/kobayashi_c.h."""
    _fields_ = [(n, C.c_double) for n in (
        "dx", "dy", "dt", "tau", "epsilon_bar", "mu", "K", "delta", "anisotropy", "alpha", "gamma", "t_eq",
        "theta0", "noise_a")]


class KobConfig(C.Structure):
    """kob_config, include/kobayashi_c.h."""
    _fields_ = [("precision", C.c_int32), ("kernel", C.c_int32), ("device", C.c_int32), ("flags", C.c_int32),
                ("seed", C.c_uint64), ("ny_global", C.c_int64), ("y0", C.c_int64)]


class KobIpcHandle(C.Structure):
    _fields_ = [("bytes", C.c_ubyte * 128)]


# name -> (restype, argtypes); every symbol declared in include/kobayashi_c.h
_P = C.c_void_p
SIGNATURES = {
    "kob_default_params": (C.c_int, [C.POINTER(KobParams), C.c_double]),
    "kob_default_config": (C.c_int, [C.POINTER(KobConfig)]),
    "kob_create": (C.c_int, [C.POINTER(_P), C.c_int64, C.c_int64, C.POINTER(KobParams), C.POINTER(KobConfig)]),
    "kob_destroy": (C.c_int, [_P]),
    "kob_reset": (C.c_int, [_P]),
    "kob_clear": (C.c_int, [_P]),
    "kob_add_nucleus": (C.c_int, [_P, C.c_int64, C.c_int64]),
    "kob_set_params": (C.c_int, [_P, C.POINTER(KobParams)]),
    "kob_get_params": (C.c_int, [_P, C.POINTER(KobParams)]),
    "kob_step": (C.c_int, [_P, C.c_int64]),
    "kob_update": (C.c_int, [_P]),
    "kob_step_timed": (C.c_int, [_P, C.c_int64, C.POINTER(C.c_float)]),
    "kob_sync": (C.c_int, [_P]),
    "kob_get_fields": (C.c_int, [_P, _P, _P, _P]),
    "kob_set_fields": (C.c_int, [_P, _P, _P, _P]),
    "kob_get_fields_async": (C.c_int, [_P, _P, _P, _P]),
    "kob_wait_fields": (C.c_int, [_P]),
    "kob_get_window": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int64, C.c_int64, _P, _P, _P]),
    "kob_set_window": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int64, C.c_int64, _P, _P, _P]),
    "kob_set_noise_field": (C.c_int, [_P, _P]),
    "kob_set_step_counter": (C.c_int, [_P, C.c_uint64]),
    "kob_get_step_counter": (C.c_int, [_P, C.POINTER(C.c_uint64)]),
    "kob_render_rgba": (C.c_int, [_P, _P]),
    "kob_sim_frame": (C.c_int, [_P, C.POINTER(C.c_int64)]),
    "kob_sim_time_ms": (C.c_int, [_P, C.POINTER(C.c_double)]),
    "kob_set_sim_counters": (C.c_int, [_P, C.c_int64, C.c_double]),
    "kob_launch_count": (C.c_int, [_P, C.POINTER(C.c_uint64)]),
    "kob_set_path_mode": (C.c_int, [_P, C.c_int32]),
    "kob_path_stats": (C.c_int, [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_double), C.POINTER(C.c_int32)]),
    "kob_concurrent_pairs": (C.c_int, [_P, C.POINTER(C.c_uint64)]),
    "kob_policy_conc_sms": (C.c_int, [C.c_int64, C.c_int64, C.c_int32, C.c_int64]),
    "kob_wait_stats": (C.c_int, [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_double)]),
    "kob_get_dims": (C.c_int, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "kob_last_error": (C.c_char_p, [_P]),
    "kob_strerror": (C.c_char_p, [C.c_int]),
    "kob_abi_version": (C.c_int, []),
    "kob_host_alloc": (C.c_int, [C.POINTER(_P), C.c_size_t]),
    "kob_host_free": (C.c_int, [_P]),
    "kob_host_alloc_near": (C.c_int, [_P, C.POINTER(_P), C.c_size_t]),
    "kob_ipc_export": (C.c_int, [_P, C.POINTER(KobIpcHandle)]),
    "kob_ipc_link": (C.c_int, [_P, C.POINTER(KobIpcHandle), C.POINTER(KobIpcHandle)]),
    "kob_link_local": (C.c_int, [_P, _P, _P]),
    "kob_halo_refresh": (C.c_int, [_P]),
    "kob_ring_join": (C.c_int, [_P, C.c_char_p, C.c_int32, C.c_int32]),
}

_lib = None


def load() -> C.CDLL:
    """Load the CUDA library; raise loudly if it is not there."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing. Build it with `python -m crystalgrowth_b200.build` "
                "(nvcc, sm_100a). crystalgrowth_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the ABI is incomplete
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib
