"""KOBCKPT1 checkpoint files (the reference has no checkpointing — Reset only re-initialises,
src/Kobayashi.cpp:241-249).  Same byte layout as Kobayashi::saveCheckpoint in include/Kobayashi.hpp:

    0   char[8]  "KOBCKPT1"
    8   u32 header bytes (256), u32 element bytes (4 | 8)
    16  i64 nx, ny, ny_global, y0          (one file per row strip)
    48  u64 step counter, u64 Philox seed, i64 sim frame
    72  f64[14] kob_params in declaration order; zero padding to 256
    256 phi, T, theta: ny*nx elements each, reference layout i + nx*j

The functions here are pure host code (numpy); `Kobayashi.save_checkpoint/load_checkpoint` move the state to and
from the device through kob_get_fields / kob_set_fields / kob_set_params / kob_set_step_counter / kob_set_sim_counters.

Not in the file: a host-injected noise field (kob_set_noise_field) — it is an input the caller owns and passes again after a
resume; the built-in Philox noise needs nothing beyond the seed and the step counter.  A ring resume is one file per strip plus
the ring-wide halo refresh (`StripRing.load_checkpoint`).
"""
from __future__ import annotations

import struct
from dataclasses import dataclass
from typing import Tuple

import numpy as np

MAGIC = b"KOBCKPT1"
HEADER_BYTES = 256
PARAM_FIELDS = ("dx", "dy", "dt", "tau", "epsilon_bar", "mu", "K", "delta", "anisotropy", "alpha", "gamma", "t_eq",
                "theta0", "noise_a")


@dataclass
class CheckpointHeader:
    elem_bytes: int
    nx: int
    ny: int
    ny_global: int
    y0: int
    step_counter: int
    seed: int
    sim_frame: int
    params: Tuple[float, ...]          # PARAM_FIELDS order

    @property
    def dtype(self):
        return np.float64 if self.elem_bytes == 8 else np.float32


def pack_header(h: CheckpointHeader) -> bytes:
    if h.elem_bytes not in (4, 8) or len(h.params) != len(PARAM_FIELDS):
        raise ValueError("bad checkpoint header")
    b = MAGIC + struct.pack("<II", HEADER_BYTES, h.elem_bytes) + struct.pack("<4q", h.nx, h.ny, h.ny_global, h.y0)
    b += struct.pack("<QQq", h.step_counter, h.seed, h.sim_frame) + struct.pack("<14d", *h.params)
    return b + bytes(HEADER_BYTES - len(b))


def unpack_header(b: bytes) -> CheckpointHeader:
    if len(b) < HEADER_BYTES or b[:8] != MAGIC:
        raise ValueError("not a KOBCKPT1 checkpoint")
    hb, eb = struct.unpack_from("<II", b, 8)
    if hb != HEADER_BYTES or eb not in (4, 8):
        raise ValueError("unsupported KOBCKPT1 header")
    nx, ny, nyg, y0 = struct.unpack_from("<4q", b, 16)
    sc, seed, fr = struct.unpack_from("<QQq", b, 48)
    return CheckpointHeader(eb, nx, ny, nyg, y0, sc, seed, fr, struct.unpack_from("<14d", b, 72))


def write_checkpoint(path: str, header: CheckpointHeader, phi, t, theta) -> None:
    shape = (header.ny, header.nx)
    with open(path, "wb") as f:
        f.write(pack_header(header))
        for a in (phi, t, theta):
            a = np.ascontiguousarray(a, header.dtype)
            if a.shape != shape:
                raise ValueError(f"field shape {a.shape} != {shape}")
            f.write(a.tobytes())


def read_checkpoint(path: str):
    """-> (header, phi, T, theta)"""
    with open(path, "rb") as f:
        h = unpack_header(f.read(HEADER_BYTES))
        n = h.nx * h.ny
        out = []
        for _ in range(3):
            raw = f.read(n * h.elem_bytes)
            if len(raw) != n * h.elem_bytes:
                raise ValueError("truncated checkpoint")
            out.append(np.frombuffer(raw, h.dtype).reshape(h.ny, h.nx).copy())
    return (h, *out)
