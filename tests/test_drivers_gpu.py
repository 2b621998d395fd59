"""The reference's UNCHANGED cuda-c sample drivers (compiled by `make drivers` from /root/reference
against include/nsparse.h + libnsparse_{s,d}.a) must run on the GPU and print their own
"Calculation Result is Correct" on the reference's only fixture (config C1)."""
import os
import re
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MTX = os.path.join(ROOT, "tests", "golden", "test.mtx")


def _run(exe, *args):
    path = os.path.join(ROOT, "bin", exe)
    if not os.path.exists(path):
        pytest.skip(f"bin/{exe} not built: the drivers are compiled from the reference tree, which is only "
                    "mounted in the build container")
    env = dict(os.environ, NSPARSE_SEED="7")
    r = subprocess.run([path, *args], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


@pytest.mark.parametrize("exe", ["spgemm_hash_d", "spgemm_hash_s"])
def test_spgemm_driver_self_check(exe):
    out = _run(exe, MTX)
    assert "Calculation Result is Correct" in out, out
    assert re.search(r"\(nnz of A\): 9 =>\s*\(Num of intermediate products\): 19 =>\s*\(nnz of C\): 11", out), out
    assert "SpGEMM using CSR format (Hash-based)" in out


@pytest.mark.parametrize("exe", ["amb_d", "amb_s"])
@pytest.mark.parametrize("plan", [(), ("2", "3"), ("65536", "1"), ("4", "20")])
def test_amb_driver_self_check(exe, plan):
    out = _run(exe, MTX, *plan)
    assert "Calculation Result is Correct" in out, out
    assert "SpMV using AMB format" in out and "Format Conversion Cost" in out
    if plan:
        assert f"(CSR=>AMB, {plan[0]}-{plan[1]})" in out, out
