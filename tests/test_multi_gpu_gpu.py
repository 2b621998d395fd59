"""N > 1 on real GPUs: every gather path of the row-blocked SpGEMM (tile pusher over NVLink peer memory, copy
kernel, copy-engine pipeline, NCCL broadcasts) against the CPU oracle and the single-GPU product, on every rank,
bit for bit.  Needs at least two GPUs in the box (the single-GPU box of the driver skips it; the host-side logic
is covered on CPU by tests/test_multi_gpu_host.py with gloo)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _ngpu():
    import torch

    return torch.cuda.device_count()


@pytest.mark.parametrize("scale", [12, 14])
def test_gather_paths_equal_oracle(scale):
    n = _ngpu()
    if n < 2:
        pytest.skip("one GPU in this box")
    n = min(n, 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "scripts", "check_mgpu.py"), "--scale", str(scale)]
    p = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:]
    assert p.stdout.count("OK") >= 2 * 5 * n and "MISMATCH" not in p.stdout, p.stdout[-3000:]
