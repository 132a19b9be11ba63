"""Host-side mirror of the reference's `Kobayashi` class (src/Kobayashi.h:32-122) on top of the C ABI.

The reference object is constructed as Kobayashi(x, y, timeStep) (src/Kobayashi.cpp:7), advanced by
iUpdate() = 10 sub-steps (src/Kobayashi.cpp:227-239), re-initialised by iResetSimulationState()
(:241-249), parameterised through its float members (_tau ... _tEq, src/Kobayashi.h:94-105) and read
through _phi (src/Kobayashi.cpp:315).  This class keeps those names and meanings; all arithmetic happens
in libkobayashi_cuda.so on the GPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import KOB_F32, KOB_F64, KOB_KERNEL_FAST, KOB_KERNEL_STRICT, KobConfig, KobIpcHandle, KobParams


class KobayashiError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"libkobayashi_cuda: {msg} (status {status})")
        self.status = status


# reference member name -> kob_params field
_PARAM_ALIASES = {
    "dx": "dx", "dy": "dy", "dt": "dt", "tau": "tau", "epsilonBar": "epsilon_bar", "epsilon_bar": "epsilon_bar",
    "mu": "mu", "K": "K", "delta": "delta", "anisotropy": "anisotropy", "alpha": "alpha", "gamma": "gamma",
    "tEq": "t_eq", "t_eq": "t_eq", "theta0": "theta0", "noise_a": "noise_a", "a": "noise_a",
}
# slider ranges of the reference GUI (src/Kobayashi.cpp:17-58): (min, max, stride)
PARAM_RANGES = {
    "tau": (0.0001, 0.0009, 0.0001), "epsilon_bar": (0.006, 0.015, 0.001), "mu": (0.5, 1.4, 0.1),
    "K": (1.0, 1.9, 0.1), "delta": (0.01, 0.09, 0.01), "anisotropy": (2.0, 8.0, 1.0),
    "alpha": (0.7, 1.2, 0.1), "gamma": (10.0, 20.0, 1.0), "t_eq": (0.5, 1.5, 0.1),
}


def default_params(dt: float = 1e-4, **over) -> KobParams:
    p = KobParams()
    _lib.load().kob_default_params(C.byref(p), dt)
    for k, v in over.items():
        setattr(p, _PARAM_ALIASES[k], v)
    return p


class Kobayashi:
    """Kobayashi(x, y, timeStep): an x-by-y periodic grid advanced with the fused CUDA step.

    precision: "f32" (what the reference stores) or "f64"; kernel: "fast" (roofline kernel, f32) or
    "strict" (reference operation order, bit-identical to the CPU oracle's portable-math build).
    ny_global / y0 make this object one row strip of a larger torus (see crystalgrowth_b200.strips).
    """

    def __init__(self, x: int, y: int, timeStep: float = 1e-4, *, precision: str = "f32", kernel: str | None = None,
                 device: int = 0, seed: int = 0, params: KobParams | None = None, ny_global: int = 0, y0: int = 0,
                 **param_overrides):
        self._L = _lib.load()
        self._h = C.c_void_p()
        prec = {"f32": KOB_F32, "f64": KOB_F64}[precision]
        if kernel is None:
            kernel = "fast" if prec == KOB_F32 else "strict"
        p = params if params is not None else default_params(timeStep)
        if params is None:
            p.dt = timeStep
        for k, v in param_overrides.items():
            setattr(p, _PARAM_ALIASES[k], v)
        cfg = KobConfig(precision=prec, kernel={"strict": KOB_KERNEL_STRICT, "fast": KOB_KERNEL_FAST}[kernel],
                        device=device, flags=0, seed=seed, ny_global=ny_global, y0=y0)
        st = self._L.kob_create(C.byref(self._h), x, y, C.byref(p), C.byref(cfg))
        if st != 0:
            msg = self._L.kob_last_error(None).decode() or self._L.kob_strerror(st).decode()
            self._h = C.c_void_p()
            raise KobayashiError(st, msg)
        self.nx, self.ny = int(x), int(y)
        self.ny_global, self.y0 = int(ny_global or y), int(y0)
        self.precision, self.kernel, self.device, self.seed = precision, kernel, device, int(seed)
        self.dtype = np.float64 if prec == KOB_F64 else np.float32

    # ---- plumbing ----
    def _ck(self, st: int):
        if st != 0:
            raise KobayashiError(st, self._L.kob_last_error(self._h).decode() or self._L.kob_strerror(st).decode())

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.kob_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def handle(self) -> C.c_void_p:
        return self._h

    # ---- reference interface ----
    def iUpdate(self):
        """10 sub-steps + frame/time accounting (src/Kobayashi.cpp:227-239)."""
        self._ck(self._L.kob_update(self._h))

    update = iUpdate

    def iResetSimulationState(self):
        """_vectorInit: zero fields, nucleus at the centre, parameters kept (src/Kobayashi.cpp:241-249)."""
        self._ck(self._L.kob_reset(self._h))

    reset = iResetSimulationState

    def step(self, n: int = 1):
        """n x {_computeGradientLaplacian; _evolution} (src/Kobayashi.cpp:230-234); asynchronous."""
        self._ck(self._L.kob_step(self._h, n))

    def step_timed(self, n: int = 1) -> float:
        ms = C.c_float()
        self._ck(self._L.kob_step_timed(self._h, n, C.byref(ms)))
        return float(ms.value)

    def sync(self):
        self._ck(self._L.kob_sync(self._h))

    def clear(self):
        self._ck(self._L.kob_clear(self._h))

    def add_nucleus(self, x: int, y: int):
        """_createNucleus(x, y) (src/Kobayashi.cpp:116-123), periodic."""
        self._ck(self._L.kob_add_nucleus(self._h, x, y))

    # parameters: read/write like the reference's members / sliders
    def get_params(self) -> KobParams:
        p = KobParams()
        self._ck(self._L.kob_get_params(self._h, C.byref(p)))
        return p

    def set_params(self, p: KobParams | None = None, *, reset: bool = False, **over):
        """Slider write (src/Kobayashi.cpp:589-611).  reset=True mirrors the GUI, which resets the fields
        after every parameter change (src/Kobayashi.cpp:616)."""
        q = p if p is not None else self.get_params()
        for k, v in over.items():
            setattr(q, _PARAM_ALIASES[k], v)
        self._ck(self._L.kob_set_params(self._h, C.byref(q)))
        if reset:
            self.reset()

    def __getattr__(self, name):
        if name in _PARAM_ALIASES and "_h" in self.__dict__:
            return getattr(self.get_params(), _PARAM_ALIASES[name])
        raise AttributeError(name)

    # ---- fields (numpy arrays of shape (ny, nx): [j, i] <-> reference index i + nx*j) ----
    def fields(self, phi=True, t=True, angl=True):
        out = [np.empty((self.ny, self.nx), self.dtype) if want else None for want in (phi, t, angl)]
        self._ck(self._L.kob_get_fields(self._h, *[None if a is None else a.ctypes.data_as(C.c_void_p) for a in out]))
        return tuple(out)

    def phi(self):
        return self.fields(True, False, False)[0]

    def t(self):
        return self.fields(False, True, False)[1]

    def angl(self):
        return self.fields(False, False, True)[2]

    def set_fields(self, phi=None, t=None, angl=None):
        arrs = []
        for a in (phi, t, angl):
            if a is None:
                arrs.append(None)
                continue
            a = np.ascontiguousarray(a, self.dtype)
            if a.shape != (self.ny, self.nx):
                raise ValueError(f"field shape {a.shape} != {(self.ny, self.nx)}")
            arrs.append(a)
        self._ck(self._L.kob_set_fields(self._h, *[None if a is None else a.ctypes.data_as(C.c_void_p) for a in arrs]))
        self.sync()

    def window(self, x0: int, y0: int, w: int, h: int, phi=True, t=True, angl=True):
        """(phi, T, theta) of the w x h window at (x0, y0): what a viewer of a huge torus reads (kob_get_window)."""
        out = [np.empty((h, w), self.dtype) if want else None for want in (phi, t, angl)]
        self._ck(self._L.kob_get_window(self._h, x0, y0, w, h, *[None if a is None else a.ctypes.data_as(C.c_void_p) for a in out]))
        return tuple(out)

    def set_window(self, x0: int, y0: int, phi=None, t=None, angl=None):
        """Write a window (all given arrays must have the same (h, w) shape); the rest of the field is kept."""
        arrs = [None if a is None else np.ascontiguousarray(a, self.dtype) for a in (phi, t, angl)]
        shapes = {a.shape for a in arrs if a is not None}
        if len(shapes) != 1:
            raise ValueError("window arrays must share one (h, w) shape")
        h, w = shapes.pop()
        self._ck(self._L.kob_set_window(self._h, x0, y0, w, h, *[None if a is None else a.ctypes.data_as(C.c_void_p) for a in arrs]))
        self.sync()

    def get_fields_into(self, phi_ptr, t_ptr, angl_ptr):
        """Raw-pointer variant (pinned host buffers) used by the end-to-end benchmark."""
        self._ck(self._L.kob_get_fields(self._h, phi_ptr, t_ptr, angl_ptr))

    def get_fields_async(self, phi_ptr, t_ptr, angl_ptr):
        """Snapshot now, copy to the (pinned) host buffers on a second stream while stepping continues; wait_fields() joins."""
        self._ck(self._L.kob_get_fields_async(self._h, phi_ptr, t_ptr, angl_ptr))

    def wait_fields(self):
        self._ck(self._L.kob_wait_fields(self._h))

    def host_alloc_near(self, nbytes: int) -> C.c_void_p:
        """Pinned host memory on the NUMA node of this context's GPU."""
        p = C.c_void_p()
        self._ck(self._L.kob_host_alloc_near(self._h, C.byref(p), nbytes))
        return p

    def set_fields_from(self, phi_ptr, t_ptr, angl_ptr):
        self._ck(self._L.kob_set_fields(self._h, phi_ptr, t_ptr, angl_ptr))

    def set_noise_field(self, r):
        if r is None:
            self._ck(self._L.kob_set_noise_field(self._h, None))
            return
        r = np.ascontiguousarray(r, np.float32)
        if r.shape != (self.ny, self.nx):
            raise ValueError("noise field shape")
        self._ck(self._L.kob_set_noise_field(self._h, r.ctypes.data_as(C.c_void_p)))

    def render_rgba(self):
        img = np.empty((self.ny, self.nx, 4), np.uint8)
        self._ck(self._L.kob_render_rgba(self._h, img.ctypes.data_as(C.c_void_p)))
        return img

    # ---- checkpoint / resume (crystalgrowth_b200.checkpoint: KOBCKPT1 files) ----
    def save_checkpoint(self, path: str):
        from . import checkpoint as ck
        p = self.get_params()
        h = ck.CheckpointHeader(np.dtype(self.dtype).itemsize, self.nx, self.ny, self.ny_global, self.y0, self.step_counter,
                                self.seed, self.simFrame, tuple(float(getattr(p, k)) for k in ck.PARAM_FIELDS))
        ck.write_checkpoint(path, h, *self.fields())

    def load_checkpoint(self, path: str):
        """Restore fields (incl. theta), parameters and the Philox step counter: the continued run is bit-identical
        to the uninterrupted one."""
        from . import checkpoint as ck
        h, phi, t, th = ck.read_checkpoint(path)
        mine = (np.dtype(self.dtype).itemsize, self.nx, self.ny, self.ny_global, self.y0, self.seed)
        if (h.elem_bytes, h.nx, h.ny, h.ny_global, h.y0, h.seed) != mine:
            raise ValueError(f"checkpoint {path} was written for another grid / precision / strip / seed")
        p = self.get_params()
        for k, v in zip(ck.PARAM_FIELDS, h.params):
            setattr(p, k, v)
        self.set_params(p)
        self.set_fields(phi, t, th)
        self.step_counter = h.step_counter
        self._ck(self._L.kob_set_sim_counters(self._h, max(0, int(h.sim_frame)), self.simTime))   # _simFrame continues
        # (a strip of a ring: follow with StripRing.refresh() / StripRing.load_checkpoint so that the neighbours see the rows)

    # ---- counters ----
    @property
    def step_counter(self) -> int:
        s = C.c_uint64()
        self._ck(self._L.kob_get_step_counter(self._h, C.byref(s)))
        return int(s.value)

    @step_counter.setter
    def step_counter(self, v: int):
        self._ck(self._L.kob_set_step_counter(self._h, v))

    @property
    def simFrame(self) -> int:
        f = C.c_int64()
        self._ck(self._L.kob_sim_frame(self._h, C.byref(f)))
        return int(f.value)

    @property
    def simTime(self) -> float:
        ms = C.c_double()
        self._ck(self._L.kob_sim_time_ms(self._h, C.byref(ms)))
        return float(ms.value)

    @property
    def launch_count(self) -> int:
        n = C.c_uint64()
        self._ck(self._L.kob_launch_count(self._h, C.byref(n)))
        return int(n.value)

    def set_path_mode(self, mode: int):
        """0 single-step kernel, 1 two-step launch pairs, 2 adaptive.  Never changes results."""
        self._ck(self._L.kob_set_path_mode(self._h, int(mode)))

    def path_stats(self) -> dict:
        """Sub-steps done by the single-step kernel / by two-step launch pairs, last density probe, policy mode."""
        a, b, fr, m = C.c_uint64(), C.c_uint64(), C.c_double(), C.c_int32()
        self._ck(self._L.kob_path_stats(self._h, C.byref(a), C.byref(b), C.byref(fr), C.byref(m)))
        n = C.c_uint64()
        self._ck(self._L.kob_concurrent_pairs(self._h, C.byref(n)))
        return {"single_steps": int(a.value), "paired_steps": int(b.value), "dense_fraction": float(fr.value), "single_mode": bool(m.value),
                "concurrent_pairs": int(n.value)}

    def wait_stats(self) -> dict:
        """Linked strips: seam-flag waits since creation and the summed waiting time of the warps concerned (ms)."""
        n, ms = C.c_uint64(), C.c_double()
        self._ck(self._L.kob_wait_stats(self._h, C.byref(n), C.byref(ms)))
        return {"waits": int(n.value), "wait_ms": float(ms.value)}

    # ---- strips ----
    def ipc_export(self) -> bytes:
        h = KobIpcHandle()
        self._ck(self._L.kob_ipc_export(self._h, C.byref(h)))
        return bytes(h.bytes)

    def ipc_link(self, lower: bytes, upper: bytes):
        hl, hu = KobIpcHandle(), KobIpcHandle()
        C.memmove(hl.bytes, lower, 128)
        C.memmove(hu.bytes, upper, 128)
        self._ck(self._L.kob_ipc_link(self._h, C.byref(hl), C.byref(hu)))

    def link_local(self, lower: "Kobayashi", upper: "Kobayashi"):
        self._ck(self._L.kob_link_local(self._h, lower._h, upper._h))

    def halo_refresh(self):
        self._ck(self._L.kob_halo_refresh(self._h))

    def ring_join(self, name: str, rank: int, world: int):
        """Ring-wide adaptive step path inside the library (kob_ring_join): every strip of the ring joins under one name."""
        self._ck(self._L.kob_ring_join(self._h, name.encode(), rank, world))
