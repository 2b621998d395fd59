"""CPU oracle for the nsparse hot paths -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package ``nsparse_b200``
never does: it fails loudly when its CUDA library is missing instead of falling back here.

``liboracle.so`` (oracle/oracle.c) restates the reference algorithms in plain C; this module is
the ctypes face of it plus numpy restatements of the comparators
(``check_spgemm_answer`` / ``ans_check``, cuda-c/src/nsparse.cu:261-353).

Parity status:
  * reader, CPU SpMV: pinned against the reference's OWN nsparse.cu compiled into
    ``oracle/_ref/libnsparse_ref_host_{s,d}.so`` (class ``ReferenceHost``), and against
    data/test.mtx's golden vectors;
  * SpGEMM: the reference has no CPU SpGEMM and its cuSPARSE oracle no longer exists, so the
    restatement is pinned by the test.mtx golden vectors and by SciPy (tests/test_oracle.py).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _build():
    subprocess.run(["make", "-C", _HERE, "liboracle.so"], check=True, stdout=subprocess.DEVNULL)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            _build()
        L = C.CDLL(path)
        L.orc_spgemm_flop.restype = C.c_longlong
        L.orc_spgemm_symbolic.restype = C.c_longlong
        L.orc_spgemm_symbolic_n.restype = C.c_longlong
        L.orc_num_threads.restype = C.c_int
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_threads(n: int) -> None:
    lib().orc_set_threads(C.c_int(n))


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def read_mtx(path: str, dtype=np.float64):
    """convert_file_csr (nsparse.cu:14-136) -> dict(M,N,nnz,nnz_max,rpt,col,val)."""
    L = lib()
    M, N, nnz, nmax = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    rpt, col, val = C.c_void_p(), C.c_void_p(), C.c_void_p()
    is_d = 1 if np.dtype(dtype) == np.float64 else 0
    rc = L.orc_mtx_read(path.encode(), is_d, C.byref(M), C.byref(N), C.byref(nnz), C.byref(nmax),
                        C.byref(rpt), C.byref(col), C.byref(val))
    if rc != 0:
        raise IOError(f"orc_mtx_read({path}) failed: {rc}")
    n = nnz.value
    out = dict(M=M.value, N=N.value, nnz=n, nnz_max=nmax.value)
    out["rpt"] = np.ctypeslib.as_array(C.cast(rpt, C.POINTER(C.c_int)), (M.value + 1,)).copy()
    out["col"] = np.ctypeslib.as_array(C.cast(col, C.POINTER(C.c_int)), (max(n, 1),))[:n].copy()
    ct = C.c_double if is_d else C.c_float
    out["val"] = np.ctypeslib.as_array(C.cast(val, C.POINTER(ct)), (max(n, 1),))[:n].copy()
    for p in (rpt, col, val):
        L.orc_free(p)
    return out


def spgemm_flop(a_rpt, a_col, b_rpt) -> int:
    a_rpt, a_col, b_rpt = _c(a_rpt, np.int32), _c(a_col, np.int32), _c(b_rpt, np.int32)
    return int(lib().orc_spgemm_flop(C.c_int(len(a_rpt) - 1), _p(a_rpt), _p(a_col), _p(b_rpt)))


def spgemm_intprod(a_rpt, a_col, b_rpt):
    a_rpt, a_col, b_rpt = _c(a_rpt, np.int32), _c(a_col, np.int32), _c(b_rpt, np.int32)
    ip = np.empty(len(a_rpt) - 1, dtype=np.int64)
    lib().orc_spgemm_intprod(C.c_int(len(a_rpt) - 1), _p(a_rpt), _p(a_col), _p(b_rpt), _p(ip))
    return ip


def spgemm_symbolic(a_rpt, a_col, b_rpt, b_col, rows=None, n_cols=0):
    """Exact row pointer (int64) of C = A*B for rows [rows[0], rows[1]).  n_cols (columns of B), when
    given, only bounds the per-row table size."""
    a_rpt, a_col = _c(a_rpt, np.int32), _c(a_col, np.int32)
    b_rpt, b_col = _c(b_rpt, np.int32), _c(b_col, np.int32)
    r0, r1 = (0, len(a_rpt) - 1) if rows is None else rows
    c_rpt = np.zeros(r1 - r0 + 1, dtype=np.int64)
    lib().orc_spgemm_symbolic_n(C.c_int(r0), C.c_int(r1), _p(a_rpt), _p(a_col), _p(b_rpt), _p(b_col), _p(c_rpt),
                                C.c_int(int(n_cols)))
    return c_rpt


def spgemm(a_rpt, a_col, a_val, b_rpt, b_col, b_val, rows=None, acc_double=True, n_cols=0):
    """C = A*B (rows [r0,r1) of A): returns (c_rpt int64, c_col int32, c_val dtype of a_val)."""
    dt = np.dtype(a_val.dtype)
    assert dt in (np.dtype(np.float32), np.dtype(np.float64))
    a_rpt, a_col, a_val = _c(a_rpt, np.int32), _c(a_col, np.int32), _c(a_val, dt)
    b_rpt, b_col, b_val = _c(b_rpt, np.int32), _c(b_col, np.int32), _c(b_val, dt)
    r0, r1 = (0, len(a_rpt) - 1) if rows is None else rows
    c_rpt = spgemm_symbolic(a_rpt, a_col, b_rpt, b_col, (r0, r1), n_cols)
    nnz = int(c_rpt[-1])
    c_col = np.empty(max(nnz, 1), dtype=np.int32)
    c_val = np.empty(max(nnz, 1), dtype=dt)
    fn = lib().orc_spgemm_numeric_d if dt == np.float64 else lib().orc_spgemm_numeric_s
    rc = fn(C.c_int(r0), C.c_int(r1), _p(a_rpt), _p(a_col), _p(a_val), _p(b_rpt), _p(b_col), _p(b_val),
            _p(c_rpt), _p(c_col), _p(c_val), C.c_int(1 if acc_double else 0))
    if rc != 0:
        raise RuntimeError("oracle numeric phase disagrees with its own symbolic phase")
    return c_rpt, c_col[:nnz], c_val[:nnz]


def spmv_csr(rpt, col, val, x, parallel=False):
    """csr_kernel (nsparse.cu:240-259)."""
    dt = np.dtype(val.dtype)
    rpt, col, val, x = _c(rpt, np.int32), _c(col, np.int32), _c(val, dt), _c(x, dt)
    y = np.empty(len(rpt) - 1, dtype=dt)
    name = "orc_spmv_csr_" + ("d" if dt == np.float64 else "s") + ("_omp" if parallel else "")
    getattr(lib(), name)(C.c_int(len(rpt) - 1), _p(rpt), _p(col), _p(val), _p(x), _p(y))
    return y


# ---- comparators -------------------------------------------------------------------------------
# north_star tolerances (BASELINE.json): column indices / nnz bit-exact, values rel 1e-6 (fp32) /
# 1e-12 (fp64).  The reference's own gates are looser (rel 1e-6 / 1e-9 for SpGEMM, 1e-5 / 1e-8 for
# SpMV, nsparse.cu:261-353).
TOL = {np.dtype(np.float32): 1e-6, np.dtype(np.float64): 1e-12}


def check_spgemm_answer(c, ans, rtol=None):
    """check_spgemm_answer (nsparse.cu:300-353): nnz, rpt, col exact; |d| <= rtol*|ans| on values.
    c / ans are (rpt, col, val).  Returns (ok, message)."""
    c_rpt, c_col, c_val = c
    a_rpt, a_col, a_val = ans
    if int(c_rpt[-1]) != int(a_rpt[-1]):
        return False, f"nnz is not correct: {int(a_rpt[-1])} (correct), {int(c_rpt[-1])} (incorrect)"
    if not np.array_equal(np.asarray(c_rpt, dtype=np.int64), np.asarray(a_rpt, dtype=np.int64)):
        i = int(np.flatnonzero(np.asarray(c_rpt, dtype=np.int64) != np.asarray(a_rpt, dtype=np.int64))[0])
        return False, f"rpt[{i}] is not correct: {a_rpt[i]} (correct), {c_rpt[i]} (incorrect)"
    if not np.array_equal(c_col, a_col):
        i = int(np.flatnonzero(c_col != a_col)[0])
        return False, f"col[{i}] is not correct: {a_col[i]} (correct), {c_col[i]} (incorrect)"
    rtol = TOL[np.dtype(a_val.dtype)] if rtol is None else rtol
    d = np.abs(a_val.astype(np.float64) - c_val.astype(np.float64))
    bad = d > rtol * np.abs(a_val.astype(np.float64))
    if bad.any():
        i = int(np.flatnonzero(bad)[0])
        return False, f"val[{i}]: ans={a_val[i]!r}, c={c_val[i]!r}, delta={d[i]:e} ({int(bad.sum())} bad)"
    return True, "Calculation Result is Correct"


def ans_check(csr_ans, ans_vec, rtol=None):
    """ans_check (nsparse.cu:261-298) with the north_star tolerance."""
    rtol = TOL[np.dtype(csr_ans.dtype)] if rtol is None else rtol
    a = csr_ans.astype(np.float64)
    d = np.abs(ans_vec.astype(np.float64) - a)
    bad = d > rtol * np.abs(ans_vec.astype(np.float64))
    if bad.any():
        i = int(np.flatnonzero(bad)[0])
        return False, f"i={i}, ans={ans_vec[i]!r}, csr={csr_ans[i]!r}, delta={d[i]:e} ({int(bad.sum())} bad)"
    return True, "Calculation Result is Correct"


# ---- the reference's own host code (oracle/_ref) -------------------------------------------------
class _sfCSR_d(C.Structure):
    _fields_ = [("rpt", C.POINTER(C.c_int)), ("col", C.POINTER(C.c_int)), ("val", C.c_void_p),
                ("d_rpt", C.c_void_p), ("d_col", C.c_void_p), ("d_val", C.c_void_p),
                ("M", C.c_int), ("N", C.c_int), ("nnz", C.c_int), ("nnz_max", C.c_int),
                ("matrix_name", C.c_char_p)]


class ReferenceHost:
    """The UNMODIFIED reference nsparse.cu (reader + csr_kernel), compiled by oracle/Makefile into
    oracle/_ref/.  C++ linkage -> mangled names."""

    def __init__(self, dtype=np.float64):
        self.dtype = np.dtype(dtype)
        suffix = "d" if self.dtype == np.float64 else "s"
        path = os.path.join(_HERE, "_ref", f"libnsparse_ref_host_{suffix}.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        self.ch = "d" if suffix == "d" else "f"

    def read_mtx(self, path: str):
        mat = _sfCSR_d()
        buf = C.create_string_buffer(path.encode())
        fn = getattr(self.lib, "_Z25init_csr_matrix_from_fileP5sfCSRPc")
        fn(C.byref(mat), buf)
        n = mat.nnz
        ct = C.c_double if self.ch == "d" else C.c_float
        out = dict(M=mat.M, N=mat.N, nnz=n, nnz_max=mat.nnz_max)
        out["rpt"] = np.ctypeslib.as_array(mat.rpt, (mat.M + 1,)).copy()
        out["col"] = np.ctypeslib.as_array(mat.col, (n,)).copy()
        out["val"] = np.ctypeslib.as_array(C.cast(mat.val, C.POINTER(ct)), (n,)).copy()
        return out

    def csr_kernel(self, rpt, col, val, x):
        rpt, col = _c(rpt, np.int32), _c(col, np.int32)
        val, x = _c(val, self.dtype), _c(x, self.dtype)
        mat = _sfCSR_d()
        mat.rpt = rpt.ctypes.data_as(C.POINTER(C.c_int))
        mat.col = col.ctypes.data_as(C.POINTER(C.c_int))
        mat.val = val.ctypes.data_as(C.c_void_p)
        mat.M = len(rpt) - 1
        mat.nnz = len(col)
        y = np.empty(len(rpt) - 1, dtype=self.dtype)
        fn = getattr(self.lib, f"_Z10csr_kernelP{self.ch}P5sfCSRS_")
        fn(_p(y), C.byref(mat), _p(x))
        return y
