"""The C-ABI library loads and exports every symbol include/nsparse_b200.h declares, the ctypes table
covers exactly those, and the nsparse.h archives export the mangled names the reference's sample
drivers import.  No compute calls: runs without a GPU."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "include", "nsparse_b200.h")
LIB = os.path.join(ROOT, "nsparse_b200", "lib", "libnsparse_b200.so")


def _declared():
    txt = open(HDR).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(nsp_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(LIB), "run `make lib` / __graft_entry__.build() first"
    L = C.CDLL(LIB)
    names = _declared()
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_ctypes_table_matches_header():
    from nsparse_b200 import _lib

    assert sorted(_lib.SIGNATURES) == _declared()
    assert not _lib._PENDING
    _lib.load()


def test_no_gpu_means_loud_failure_not_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import nsparse_b200 as ns

    with pytest.raises(ns.NsparseError):
        ns.Context(0)
    L = ns.load_library()
    h = C.c_void_p()
    assert L.nsp_create(C.byref(h), 0) != 0 and not h.value


def test_product_does_not_reference_the_oracle():
    for dp, _, files in os.walk(os.path.join(ROOT, "nsparse_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f


MANGLED = ["_Z25init_csr_matrix_from_fileP5sfCSRPc", "_Z10csr_memcpyP5sfCSR", "_Z13csr_memcpyDtHP5sfCSR",
           "_Z11release_csr5sfCSR", "_Z15release_cpu_csr5sfCSR", "_Z15get_spgemm_flopP5sfCSRS0_iPx",
           "_Z18spgemm_kernel_hashP5sfCSRS0_S0_", "_Z13spgemm_cu_csrP5sfCSRS0_S0_",
           "_Z19check_spgemm_answer5sfCSRS_", "_Z9init_planP6sfPlan", "_Z8set_planP6sfPlanmi",
           "_Z11release_amb5sfAMB"]
PER_PREC = ["_Z11init_vectorP{t}i", "_Z10csr_kernelP{t}P5sfCSRS_", "_Z10sf_csr2ambP5sfAMBP5sfCSRP{t}P6sfPlan",
            "_Z11sf_spmv_ambP{t}P5sfAMBS_P6sfPlan", "_Z9ans_checkP{t}S_i"]


@pytest.mark.parametrize("prec,t", [("s", "f"), ("d", "d")])
def test_compat_archives_export_the_driver_symbols(prec, t):
    ar = os.path.join(ROOT, "nsparse_b200", "lib", f"libnsparse_{prec}.a")
    if not os.path.exists(ar):
        pytest.skip("compat archives not built (make compat)")
    out = subprocess.run(["nm", "--defined-only", ar], capture_output=True, text=True).stdout
    defined = set(re.findall(r" T (\S+)", out))
    want = MANGLED + [s.format(t=t) for s in PER_PREC]
    missing = [s for s in want if s not in defined]
    assert not missing, missing
