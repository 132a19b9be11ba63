"""pytest configuration: the `gpu` marker, import path, and shared helpers.

CPU suite (`-m "not gpu"`): oracle vs reference TU / golden vectors, host logic, C-ABI surface, gloo strips.
GPU suite (`-m gpu`): parity of libkobayashi_cuda.so (through the C ABI) against the oracle.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def bit_equal(a, b) -> bool:
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    return a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a.view(np.uint8), b.view(np.uint8))


def max_abs(a, b) -> float:
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max())


@pytest.fixture(scope="session")
def po():
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def cg():
    """The product package.  Fails (does not skip) when the CUDA library is missing: no CPU fallback exists."""
    import crystalgrowth_b200
    crystalgrowth_b200.load()
    return crystalgrowth_b200
