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


@pytest.mark.parametrize("exe", ["spgemm_cu_csr_d", "spgemm_cu_csr_s"])
def test_cusparse_spgemm_driver(exe):
    """f3: the reference's UNCHANGED cuSPARSE comparison driver (sample/spgemm/spgemm_cu_csr.cu) against
    spgemm_kernel_cu_csr on the generic cusparseSpGEMM API."""
    out = _run(exe, MTX)
    assert "SpGEMM using CSR format (cuSPARSE)" in out, out
    assert re.search(r"\(nnz of A\): 9 =>\s*\(Num of intermediate products\): 19 =>\s*\(nnz of C\): 11", out), out


@pytest.mark.parametrize("exe", ["cu_csr_d", "cu_csr_s"])
def test_cusparse_spmv_driver(exe):
    """f3: cuSPARSE CSR SpMV comparison (sf_spmv_cu_csr on cusparseSpMV; the reference's driver calls the removed
    legacy csrmv and cannot be compiled unchanged)."""
    out = _run(exe, MTX)
    assert "Calculation Result is Correct" in out, out
    assert "SpMV using CSR format (cuSPARSE)" in out


def _write_mtx(path, a):
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n")
        f.write(f"{a.M} {a.N} {a.nnz}\n")
        import numpy as np

        rows = np.repeat(np.arange(a.M), np.diff(a.rpt))
        for r, c, v in zip(rows, a.col, a.val):
            f.write(f"{r + 1} {c + 1} {v:.9g}\n")


@pytest.mark.parametrize("exe", ["spgemm_hash_mgpu_d", "spgemm_hash_mgpu_s"])
def test_mgpu_driver_self_check(exe, tmp_path):
    """The C/C++ multi-GPU entry (spgemm_kernel_hash_mgpu, one process, all GPUs of the box; with one GPU it
    degenerates to a single block): every GPU's copy of C passes check_spgemm_answer against the single-GPU
    product, on data/test.mtx and on an R-MAT matrix that reaches the heavy classes."""
    import numpy as np

    from nsparse_b200 import gen

    out = _run(exe, MTX)
    assert "Calculation Result is Incorrect" not in out and out.count("Calculation Result is Correct") >= 1, out
    a = gen.rmat_csr(12, 16, seed=11, dtype=np.float64, values="small_int")
    p = str(tmp_path / "rmat12.mtx")
    _write_mtx(p, a)
    out = _run(exe, p)
    assert "Calculation Result is Incorrect" not in out and out.count("Calculation Result is Correct") >= 1, out
    assert "Hash-based," in out
