"""CPU check of the FAST kernels' data-dependent arithmetic (no GPU needed).

crystalgrowth_b200/csrc/kob_row.cuh's cold_block() — the re-assigned angle, the trig-free anisotropy (closed forms for
j = 4, 6; repeated squaring for other integer j; rotation by theta0; trig for real j), m(T) and the noise term — is written
against primitives that also compile for the host.  tests/cpp/cold_check.cpp compiles it with g++ (-DKOB_HOST_EMU) and
compares ~3 M random cells, including dead-band, held-angle, +-0 and diagonal cases, with the reference's own expressions
(src/Kobayashi.cpp:154-171, :206-214) evaluated with libm.  It pins the formulas before any GPU time is spent; the GPU
parity tests (tests/test_gpu_parity.py) then hold the compiled kernels to the 1e-6 single-step tolerance."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cold_block_matches_reference_expressions(tmp_path):
    exe = tmp_path / "cold_check"
    src = os.path.join(ROOT, "tests", "cpp", "cold_check.cpp")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-std=c++17", "-O2", "-DKOB_HOST_EMU", "-ffp-contract=off", "-o", str(exe), src], check=True)
    r = subprocess.run([str(exe)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout[-2000:]
