"""crystalgrowth_b200 — B200-native (sm_100a) implementation of CrystalGrowth's hot path: the fused
explicit-Euler step of Kobayashi's anisotropic phase-field model, behind the reference's `Kobayashi`
class interface.  All compute lives in libkobayashi_cuda.so (C ABI: include/kobayashi_c.h)."""
from ._lib import KOB_F32, KOB_F64, KOB_KERNEL_FAST, KOB_KERNEL_STRICT, KobConfig, KobParams, load  # noqa: F401
from .kobayashi import PARAM_RANGES, Kobayashi, KobayashiError, default_params  # noqa: F401

__all__ = ["Kobayashi", "KobayashiError", "KobParams", "KobConfig", "default_params", "PARAM_RANGES", "load"]
