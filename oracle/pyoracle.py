"""ctypes bindings for the CPU oracle (oracle/_build/libkob_oracle.so) and for the reference translation
unit compiled in place (oracle/_ref/libkobref_f{32,64}.so).

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product package crystalgrowth_b200 never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "libkob_oracle.so")
REF_SO = {32: os.path.join(HERE, "_ref", "libkobref_f32.so"), 64: os.path.join(HERE, "_ref", "libkobref_f64.so")}

MATH_LIBM, MATH_PORTABLE = 0, 1


class KobParams(C.Structure):
    """Mirror of kob_params (include/kobayashi_c.h)."""
    _fields_ = [(n, C.c_double) for n in (
        "dx", "dy", "dt", "tau", "epsilon_bar", "mu", "K", "delta", "anisotropy", "alpha", "gamma", "t_eq",
        "theta0", "noise_a")]


def default_params(dt: float = 1e-4, **over) -> KobParams:
    """Reference defaults, src/Kobayashi.cpp:61-63 and :76-84."""
    p = KobParams(dx=0.03, dy=0.03, dt=dt, tau=0.0003, epsilon_bar=0.010, mu=1.0, K=1.6, delta=0.05,
                  anisotropy=6.0, alpha=0.9, gamma=10.0, t_eq=1.0, theta0=0.0, noise_a=0.0)
    for k, v in over.items():
        if not hasattr(p, k):
            raise KeyError(k)
        setattr(p, k, v)
    return p


def float_rounded(p: KobParams) -> KobParams:
    """Parameters as the FP64-typed reference build sees them: float literals widened to double."""
    q = KobParams()
    for name, _ in KobParams._fields_:
        setattr(q, name, float(np.float32(getattr(p, name))))
    return q


def build(force: bool = False) -> None:
    """make -C oracle (restatement always; oracle/_ref only where /root/reference exists)."""
    if force or not os.path.exists(ORACLE_SO) or (os.path.exists("/root/reference/src/Kobayashi.cpp")
                                                 and not os.path.exists(REF_SO[32])):
        subprocess.run(["make", "-C", HERE, "all"], check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(ORACLE_SO)
        L.kobo_create.restype = C.c_void_p
        L.kobo_create.argtypes = [C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.POINTER(KobParams),
                                  C.c_int, C.c_uint64]
        for name, args in {
            "kobo_destroy": [C.c_void_p], "kobo_clear": [C.c_void_p], "kobo_reset": [C.c_void_p],
            "kobo_add_nucleus": [C.c_void_p, C.c_int64, C.c_int64],
            "kobo_set_params": [C.c_void_p, C.POINTER(KobParams)],
            "kobo_step": [C.c_void_p, C.c_int64],
            "kobo_get_fields": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
            "kobo_set_fields": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
            "kobo_get_edge": [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p],
            "kobo_set_ghost": [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p],
            "kobo_set_noise_field": [C.c_void_p, C.c_void_p],
            "kobo_set_step_counter": [C.c_void_p, C.c_uint64],
            "kobo_set_threads": [C.c_void_p, C.c_int],
            "kobo_set_noise_origin": [C.c_void_p, C.c_int64, C.c_int64],
            "kobo_philox": [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)],
        }.items():
            getattr(L, name).restype = None
            getattr(L, name).argtypes = args
        L.kobo_get_step_counter.restype = C.c_uint64
        L.kobo_get_step_counter.argtypes = [C.c_void_p]
        L.kobo_max_threads.restype = C.c_int
        L.kobo_noise_r.restype = C.c_float
        L.kobo_noise_r.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32]
        for n, t in (("kobo_p_atanf", C.c_float), ("kobo_p_sinf", C.c_float), ("kobo_p_cosf", C.c_float),
                     ("kobo_p_atan", C.c_double), ("kobo_p_sin", C.c_double), ("kobo_p_cos", C.c_double)):
            getattr(L, n).restype = t
            getattr(L, n).argtypes = [t]
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    """CPU restatement of the Kobayashi step.  Fields are numpy arrays of shape (ny, nx) — row j, column i —
    which is the reference layout i + nx*j."""

    def __init__(self, nx, ny, params=None, prec=32, math=MATH_LIBM, seed=0, ny_global=0, y0=0, threads=1,
                 reset=True):
        self.nx, self.ny, self.prec = int(nx), int(ny), int(prec)
        self.dtype = np.float64 if prec == 64 else np.float32
        self.params = params if params is not None else default_params()
        self._h = lib().kobo_create(prec, nx, ny, ny_global, y0, C.byref(self.params), math, seed)
        if not self._h:
            raise MemoryError("kobo_create failed")
        lib().kobo_set_threads(self._h, threads)
        if reset:
            self.reset()

    def __del__(self):
        if getattr(self, "_h", None):
            lib().kobo_destroy(self._h)
            self._h = None

    def reset(self): lib().kobo_reset(self._h)
    def clear(self): lib().kobo_clear(self._h)
    def add_nucleus(self, x, y): lib().kobo_add_nucleus(self._h, x, y)
    def step(self, n=1): lib().kobo_step(self._h, n)
    def set_threads(self, n): lib().kobo_set_threads(self._h, n)
    def set_noise_origin(self, x0, y0): lib().kobo_set_noise_origin(self._h, int(x0), int(y0))

    def set_params(self, params):
        self.params = params
        lib().kobo_set_params(self._h, C.byref(params))

    def fields(self):
        phi = np.empty((self.ny, self.nx), self.dtype)
        t = np.empty_like(phi)
        a = np.empty_like(phi)
        lib().kobo_get_fields(self._h, _ptr(phi), _ptr(t), _ptr(a))
        return phi, t, a

    def set_fields(self, phi=None, t=None, angl=None):
        arrs = [None if x is None else np.ascontiguousarray(x, self.dtype) for x in (phi, t, angl)]
        for x in arrs:
            assert x is None or x.shape == (self.ny, self.nx)
        lib().kobo_set_fields(self._h, *[_ptr(x) for x in arrs])

    def edge(self, side):
        out = [np.empty((2, self.nx), self.dtype) for _ in range(3)]
        lib().kobo_get_edge(self._h, side, *[_ptr(x) for x in out])
        return out

    def set_ghost(self, side, phi, t, angl):
        arrs = [np.ascontiguousarray(x, self.dtype) for x in (phi, t, angl)]
        lib().kobo_set_ghost(self._h, side, *[_ptr(x) for x in arrs])

    def set_noise_field(self, r):
        if r is None:
            lib().kobo_set_noise_field(self._h, None)
        else:
            r = np.ascontiguousarray(r, np.float32)
            assert r.shape == (self.ny, self.nx)
            lib().kobo_set_noise_field(self._h, _ptr(r))

    def set_step_counter(self, s): lib().kobo_set_step_counter(self._h, s)
    def step_counter(self): return int(lib().kobo_get_step_counter(self._h))


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().kobo_philox(c, k, o)
    return [int(x) for x in o]


def noise_r(seed, step, i, j):
    return float(lib().kobo_noise_r(seed, step, i, j))


# ------------------------------------------------------------------------------------------------------
# The reference translation unit itself (only where oracle/_ref was built or shipped).
# ------------------------------------------------------------------------------------------------------
PARAM_INDEX = {"dx": 0, "dy": 1, "dt": 2, "tau": 3, "epsilon_bar": 4, "mu": 5, "K": 6, "delta": 7,
               "anisotropy": 8, "alpha": 9, "gamma": 10, "t_eq": 11}
_ref_libs = {}


def ref_available(prec=32) -> bool:
    build()
    return os.path.exists(REF_SO[prec])


def ref_lib(prec=32) -> C.CDLL:
    if prec not in _ref_libs:
        if not ref_available(prec):
            raise FileNotFoundError(REF_SO[prec])
        L = C.CDLL(REF_SO[prec])
        L.ref_create.restype = C.c_void_p
        L.ref_create.argtypes = [C.c_int, C.c_int, C.c_double]
        L.ref_get_param.restype = C.c_double
        L.ref_get_param.argtypes = [C.c_void_p, C.c_int]
        for name, args in {
            "ref_destroy": [C.c_void_p], "ref_reset": [C.c_void_p], "ref_update": [C.c_void_p],
            "ref_set_param": [C.c_void_p, C.c_int, C.c_double], "ref_step": [C.c_void_p, C.c_int64],
            "ref_get_fields": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
            "ref_set_fields": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
            "ref_add_nucleus": [C.c_void_p, C.c_int, C.c_int], "ref_colors": [C.c_void_p, C.c_void_p],
            "ref_gui_attach": [C.c_void_p], "ref_gui_command": [C.c_void_p, C.c_int],
            "ref_gui_hscroll": [C.c_void_p, C.c_int, C.c_int, C.c_int], "ref_gui_frame": [C.c_void_p],
            "ref_gui_colors": [C.c_void_p, C.c_void_p], "ref_gui_state": [C.c_void_p, C.c_void_p],
        }.items():
            getattr(L, name).restype = None
            getattr(L, name).argtypes = args
        _ref_libs[prec] = L
    return _ref_libs[prec]


class Reference:
    """The unmodified reference class `Kobayashi` (src/Kobayashi.h:32) driven through ref_driver.cpp."""

    def __init__(self, nx, ny, dt=1e-4, prec=32, **params):
        self.nx, self.ny, self.prec = int(nx), int(ny), prec
        self.dtype = np.float64 if prec == 64 else np.float32
        self._L = ref_lib(prec)
        self._h = self._L.ref_create(nx, ny, dt)
        for k, v in params.items():
            self.set_param(k, v)

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.ref_destroy(self._h)
            self._h = None

    def set_param(self, name, v): self._L.ref_set_param(self._h, PARAM_INDEX[name], float(v))
    def get_param(self, name): return float(self._L.ref_get_param(self._h, PARAM_INDEX[name]))
    def reset(self): self._L.ref_reset(self._h)
    def step(self, n=1): self._L.ref_step(self._h, n)
    def update(self): self._L.ref_update(self._h)
    def add_nucleus(self, x, y): self._L.ref_add_nucleus(self._h, x, y)

    def fields(self):
        phi = np.empty((self.ny, self.nx), self.dtype)
        t = np.empty_like(phi)
        a = np.empty_like(phi)
        self._L.ref_get_fields(self._h, _ptr(phi), _ptr(t), _ptr(a))
        return phi, t, a

    def set_fields(self, phi=None, t=None, angl=None):
        arrs = [None if x is None else np.ascontiguousarray(x, self.dtype) for x in (phi, t, angl)]
        self._L.ref_set_fields(self._h, *[_ptr(x) for x in arrs])

    def colors(self):
        rgb = np.empty((self.nx * self.ny, 3), self.dtype)
        self._L.ref_colors(self._h, _ptr(rgb))
        return rgb

    # ---- the reference's control panel, driven headlessly (src/Kobayashi.cpp:383-629) ----
    def gui_attach(self): self._L.ref_gui_attach(self._h)
    def gui_command(self, com): self._L.ref_gui_command(self._h, int(com))
    def gui_hscroll(self, index, code, pos=0): self._L.ref_gui_hscroll(self._h, int(index), int(code), int(pos))
    def gui_frame(self): self._L.ref_gui_frame(self._h)

    def gui_colors(self):
        rgb = np.zeros((self.nx * self.ny, 3), self.dtype)
        self._L.ref_gui_colors(self._h, _ptr(rgb))
        return rgb

    def gui_state(self):
        out = np.zeros(20, np.float64)
        self._L.ref_gui_state(self._h, _ptr(out))
        return {"playing": bool(out[0]), "sim_frame": int(out[1]), "values": out[2:11].copy(), "positions": out[11:20].astype(int)}
