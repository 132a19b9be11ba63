"""The headless C++ driver (kob_bench) and the C++ host class Kobayashi.hpp: the reference's default run
(src/main.cpp:14-18: 250x250, dt = 1e-4; 200 iUpdate calls = 2000 sub-steps) through the C ABI from C++."""
import json
import os
import subprocess

import pytest

from conftest import ROOT

DRIVER = os.path.join(ROOT, "crystalgrowth_b200", "driver", "kob_bench")


def test_driver_builds_and_links_against_the_c_abi():
    from crystalgrowth_b200 import build
    build.build_all()
    assert os.path.exists(DRIVER)
    out = subprocess.run(["ldd", DRIVER], stdout=subprocess.PIPE, text=True).stdout
    assert "libkobayashi_cuda.so" in out and "not found" not in out.split("libkobayashi_cuda.so")[1].split("\n")[0]


def test_cpp_header_compiles_standalone():
    src = '#include "Kobayashi.hpp"\nint main() { return sizeof(Kobayashi) > 0 ? 0 : 1; }\n'
    r = subprocess.run(["g++", "-std=c++14", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), "-x", "c++", "-"],
                       input=src, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("kernel", ["fast", "strict"])
def test_reference_default_run_headless(kernel):
    r = subprocess.run([DRIVER, "--nx", "250", "--ny", "250", "--frames", "199", "--kernel", kernel],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["sim_frames"] == 200 and d["launches"] >= 2000
    assert abs(d["solid_cells"] / 10817 - 1) <= 0.02       # 10817 solid cells after 2000 sub-steps
    assert d["gcell_per_s_device"] > 0
