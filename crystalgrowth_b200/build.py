"""In-tree build of libkobayashi_cuda.so (nvcc, sm_100a only) and of the headless C++ driver kob_bench.

    python -m crystalgrowth_b200.build [--force]

nvcc cross-compiles without a GPU.  Outputs stay in-tree (git-ignored) so that they travel to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libkobayashi_cuda.so")
BENCH = os.path.join(PKG, "driver", "kob_bench")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-ffp-contract=off", "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libkobayashi_cuda.so cannot be built (there is no CPU fallback)")


def _host_cxx() -> str:
    # the image exports CXX=/opt/gcc/bin/g++ (wrapper); the distro compiler is what nvcc was validated with
    return "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def _stale(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def lib_sources() -> list[str]:
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".h"))]
    srcs.append(os.path.join(ROOT, "include", "kobayashi_c.h"))
    return srcs


def build_lib(force: bool = False, verbose: bool = False) -> str:
    if force or _stale(LIB, lib_sources()):
        cmd = [_nvcc(), *NVCC_FLAGS, "-ccbin", _host_cxx(), "-shared", "-o", LIB, os.path.join(CSRC, "kob_api.cu")]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + r.stdout)
        if verbose:
            print(r.stdout)
    return LIB


def build_variant(name: str, defines: list[str]) -> str:
    """Development: the same library with other compile-time knobs (-DKOB_FAST_WARPS=12 ...), loaded through
    KOB_LIB_PATH (see _lib.py).  Output: crystalgrowth_b200/variants/libkobayashi_cuda_<name>.so (git-ignored)."""
    out_dir = os.path.join(PKG, "variants")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, f"libkobayashi_cuda_{name}.so")
    cmd = [_nvcc(), *NVCC_FLAGS, *defines, "-Xptxas=-v", "-ccbin", _host_cxx(), "-shared", "-o", out, os.path.join(CSRC, "kob_api.cu")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout)
    with open(out + ".log", "w") as fh:
        fh.write(r.stdout)
    return out


def build_driver(force: bool = False) -> str:
    src = os.path.join(PKG, "driver", "kob_bench.cpp")
    hdrs = [src, os.path.join(ROOT, "include", "kobayashi_c.h"), os.path.join(ROOT, "include", "Kobayashi.hpp")]
    if os.path.exists(src) and (force or _stale(BENCH, hdrs)):
        cmd = [_host_cxx(), "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"), "-o", BENCH, src,
               "-L", PKG, "-lkobayashi_cuda", "-Wl,-rpath,$ORIGIN/..", "-ldl"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("driver build failed:\n" + r.stdout)
    return BENCH


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_lib(force, verbose)
    build_driver(force)


if __name__ == "__main__":
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], [a for a in sys.argv[i + 2:] if a.startswith("-D")]))
    else:
        build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
        print(LIB)
