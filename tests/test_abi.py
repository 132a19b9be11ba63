"""The C-ABI surface: libkobayashi_cuda.so loads without a GPU and exports every symbol that
include/kobayashi_c.h declares; without a device the entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "kobayashi_c.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(kob_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    names = declared_symbols()
    for must in ("kob_create", "kob_destroy", "kob_reset", "kob_add_nucleus", "kob_set_params", "kob_step", "kob_update",
                 "kob_get_fields", "kob_set_fields", "kob_set_noise_field", "kob_ipc_export", "kob_ipc_link"):
        assert must in names


def test_library_exports_every_declared_symbol(cg):
    lib = cg.load()
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in kobayashi_c.h but not exported"
    # and the Python signature table covers the header exactly
    from crystalgrowth_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    assert lib.kob_abi_version() == 1


def test_header_is_plain_c():
    """The boundary must be consumable from C (cgo / JNI / ctypes style FFI): compile it with gcc -std=c99."""
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", HEADER],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout


def test_struct_layouts_match(cg):
    from crystalgrowth_b200._lib import KobConfig, KobIpcHandle, KobParams
    assert C.sizeof(KobParams) == 14 * 8
    assert C.sizeof(KobConfig) == 4 * 4 + 8 + 8 + 8
    assert C.sizeof(KobIpcHandle) == 128


def test_defaults_are_the_reference_defaults(cg):
    p = cg.default_params(1e-4)     # src/Kobayashi.cpp:61-63, :76-84
    assert (p.dx, p.dy, p.dt, p.tau, p.epsilon_bar, p.mu, p.K, p.delta, p.anisotropy, p.alpha, p.gamma, p.t_eq) == \
        (0.03, 0.03, 1e-4, 0.0003, 0.010, 1.0, 1.6, 0.05, 6.0, 0.9, 10.0, 1.0)
    assert p.theta0 == 0.0 and p.noise_a == 0.0


def test_no_cpu_fallback_without_device(cg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(cg.KobayashiError) as e:
        cg.Kobayashi(32, 32, 1e-4)
    assert e.value.status == -3     # KOB_ERR_NO_DEVICE


def test_product_does_not_import_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "crystalgrowth_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "pyoracle" not in txt and "kob_oracle" not in txt and "libkob_oracle" not in txt, f


def test_launch_pair_policy_is_a_pure_function(cg):
    """How many SMs the general pass of a launch pair gets beside its far pass, without a device: grows with the
    work list, never exceeds 40 % of the SMs, falls back to the plain far -> general order (0) for long lists and tiny devices."""
    lib = cg.load()
    f = lambda est, n=16384, sms=148: lib.kob_policy_conc_sms(n, n, sms, est)      # noqa: E731
    g = [f(e) for e in (0, 100, 400, 800, 1500, 2300, 3100)]
    assert g[0] == 2 and all(b >= a for a, b in zip(g, g[1:])), g
    assert 18 <= f(1500) <= 26 and 36 <= f(3100) <= 50, g                          # the measured sweet spots (profiles/r02_bench_summary.md)
    assert max(g) <= 148 * 2 // 5
    assert f(8000) == 0 and f(10 ** 7) == 0                                        # a listed fraction beyond ~3 %: plain order
    assert f(100, sms=8) == 0                                                       # not worth it on a sliver of a device
    assert f(-1) == 0                                                               # no probe yet
    # a small grid: the far pass is short, a ticket does not fit beside it more than once per warp
    assert f(50, n=4096) >= 1 and f(5000, n=4096) == 0
