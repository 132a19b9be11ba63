"""G5 on real hardware: P row strips on P GPUs (one process each, CUDA-IPC ring) == the single-GPU run, bitwise.
Needs >= 2 visible GPUs; skipped otherwise (the 1-GPU box covers the same kernel paths through
test_linked_strips_on_one_gpu_equal_single_domain)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.timeout(600)
@pytest.mark.parametrize("kernel,precision", [("fast", "f32"), ("strict", "f32"), ("strict", "f64")])
def test_ipc_ring_equals_single_gpu(kernel, precision):
    n = _ngpu()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 4 if n >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "scripts", "mgpu_check.py"), "--kernel", kernel, "--precision", precision]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=550)
    assert r.returncode == 0, r.stdout[-3000:]
    assert r.stdout.count("bitwise=True") == 3, r.stdout[-3000:]
