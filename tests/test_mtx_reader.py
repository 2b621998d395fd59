"""nsp_read_mtx (csrc/mtx_reader.cpp): the parallel MatrixMarket reader must give, entry
for entry, what the reference reader convert_file_csr (nsparse.cu:14-136) gives -- checked against the committed
reader goldens, against the CPU oracle's restatement on generated files, and (when the reference tree is mounted
and `make -C oracle ref` was run) against the reference's own compiled reader.  No GPU needed."""
import ctypes as C
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp

from nsparse_b200 import _lib
from oracle import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = json.load(open(os.path.join(HERE, "golden", "reader_cases.json")))


def read(path, dtype=np.float64, flags=0):
    L = _lib.load()
    M, N, nmax, nnz = C.c_int(), C.c_int(), C.c_int(), C.c_longlong()
    rpt, col, val = C.c_void_p(), C.c_void_p(), C.c_void_p()
    is_d = 1 if np.dtype(dtype) == np.float64 else 0
    rc = L.nsp_read_mtx(str(path).encode(), is_d, flags, C.byref(M), C.byref(N), C.byref(nnz), C.byref(nmax),
                        C.byref(rpt), C.byref(col), C.byref(val))
    if rc != 0:
        raise IOError(rc)
    n = nnz.value
    out = dict(M=M.value, N=N.value, nnz=n, nnz_max=nmax.value)
    out["rpt"] = np.ctypeslib.as_array(C.cast(rpt, C.POINTER(C.c_int)), (M.value + 1,)).copy()
    out["col"] = np.ctypeslib.as_array(C.cast(col, C.POINTER(C.c_int)), (max(n, 1),))[:n].copy()
    ct = C.c_double if is_d else C.c_float
    out["val"] = np.ctypeslib.as_array(C.cast(val, C.POINTER(ct)), (max(n, 1),))[:n].copy()
    for p in (rpt, col, val):
        L.nsp_free_host(p)
    return out


def same(a, b):
    return (a["M"], a["N"], a["nnz"], a["nnz_max"]) == (b["M"], b["N"], b["nnz"], b["nnz_max"]) and \
        np.array_equal(a["rpt"], b["rpt"]) and np.array_equal(a["col"], b["col"]) and np.array_equal(a["val"], b["val"])


@pytest.mark.parametrize("name", sorted(CASES))
def test_reader_goldens(tmp_path, name):
    case = CASES[name]
    f = tmp_path / "m.mtx"
    f.write_text(case["text"])
    got = read(f)
    assert (got["M"], got["N"], got["nnz"], got["nnz_max"]) == (case["M"], case["N"], case["nnz"], case["nnz_max"])
    assert got["rpt"].tolist() == case["rpt"] and got["col"].tolist() == case["col"] and got["val"].tolist() == case["val"]


def _write_random(path, m, n, nz, symmetric, pattern, seed):
    rng = np.random.default_rng(seed)
    r = rng.integers(1, m + 1, size=nz)
    c = rng.integers(1, n + 1, size=nz)
    if symmetric:
        r, c = np.maximum(r, c), np.minimum(r, c)
    v = rng.standard_normal(nz)
    kind = "pattern" if pattern else "real"
    with open(path, "w") as f:
        f.write(f"%%MatrixMarket matrix coordinate {kind} {'symmetric' if symmetric else 'general'}\n% comment\n%\n")
        f.write(f"{m} {n} {nz}\n")
        for i in range(nz):
            f.write(f"{r[i]} {c[i]}\n" if pattern else f"{r[i]} {c[i]} {float(v[i])!r}\n")


@pytest.mark.parametrize("symmetric,pattern", [(False, False), (True, False), (True, True), (False, True)])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_reader_matches_oracle_on_large_files(tmp_path, symmetric, pattern, dtype):
    """300 k lines: big enough for the multi-threaded path (pieces cut at line boundaries, concatenated in
    order), duplicates and unsorted rows included."""
    f = tmp_path / "big.mtx"
    _write_random(f, 5000, 5000 if symmetric else 7000, 300_000, symmetric, pattern, seed=11)
    assert same(read(f, dtype), oracle.read_mtx(str(f), dtype))


def test_reader_against_reference_binary(tmp_path):
    try:
        ref = oracle.ReferenceHost(np.float64)
    except (FileNotFoundError, OSError):
        pytest.skip("oracle/_ref not built (reference tree not mounted)")
    f = tmp_path / "sym.mtx"
    _write_random(f, 3000, 3000, 100_000, True, False, seed=5)
    assert same(read(f, np.float64), ref.read_mtx(str(f)))


def test_sort_merge_option(tmp_path):
    f = tmp_path / "dup.mtx"
    _write_random(f, 400, 300, 20_000, False, False, seed=3)     # many duplicates
    got = read(f, np.float64, flags=1)
    raw = oracle.read_mtx(str(f), np.float64)
    want = sp.csr_matrix((raw["val"], raw["col"], raw["rpt"]), shape=(raw["M"], raw["N"]))
    want.sum_duplicates()
    want.sort_indices()
    assert got["nnz"] == want.nnz and np.array_equal(got["rpt"], want.indptr) and np.array_equal(got["col"], want.indices)
    np.testing.assert_allclose(got["val"], want.data, rtol=1e-12, atol=1e-12)
    assert np.all(np.diff(got["col"])[np.setdiff1d(np.arange(got["nnz"] - 1), got["rpt"][1:-1] - 1)] > 0)


def test_unparsable_value_reads_as_zero_like_atof(tmp_path):
    f = tmp_path / "odd.mtx"
    f.write_text("%%MatrixMarket matrix coordinate real general\n2 2 3\n1 1 abc\n2 1\n2 2 1e2\n")
    got = read(f)
    assert same(got, oracle.read_mtx(str(f), np.float64)) and got["val"].tolist() == [0.0, 1.0, 100.0]


def test_tabs_crlf_blank_lines_and_entry_limit(tmp_path):
    """White space variants the reference's strchr(' ') tokenizer never sees are read the natural way; lines
    beyond the nz of the size line are ignored like the reference's fixed-size COO arrays would require."""
    f = tmp_path / "ws.mtx"
    f.write_bytes(b"%%MatrixMarket matrix coordinate real general\r\n% c\r\n3 3 3\r\n1\t2\t2.5\r\n\r\n  3 1 -1\r\n2 2 7\r\n1 1 99\r\n")
    got = read(f)
    assert (got["M"], got["N"], got["nnz"]) == (3, 3, 3)
    assert got["rpt"].tolist() == [0, 1, 2, 3] and got["col"].tolist() == [1, 1, 0] and got["val"].tolist() == [2.5, 7.0, -1.0]


def test_reader_errors(tmp_path):
    with pytest.raises(IOError):
        read(tmp_path / "missing.mtx")
    f = tmp_path / "bad.mtx"
    f.write_text("%%MatrixMarket matrix coordinate real general\n2 2 1\n3 1 1.0\n")
    with pytest.raises(IOError):
        read(f)
