"""CPU tests of the oracle: golden vectors of data/test.mtx, the reference's own host code
(oracle/_ref), and SciPy.  No GPU."""
import json
import os
import tempfile

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "test_mtx.json")))
CASES = json.load(open(os.path.join(HERE, "golden", "reader_cases.json")))
MTX = os.path.join(HERE, "golden", "test.mtx")


def _rand_csr(m, n, density, seed, dtype=np.float64, sort=True):
    rng = np.random.default_rng(seed)
    a = sp.random(m, n, density=density, random_state=rng, format="csr", dtype=np.float64)
    a.data = rng.integers(1, 4, size=a.nnz).astype(dtype)
    if sort:
        a.sort_indices()
    return a


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_reader_golden(dtype):
    a = oracle.read_mtx(MTX, dtype)
    assert (a["M"], a["N"], a["nnz"], a["nnz_max"]) == (GOLD["M"], GOLD["N"], GOLD["nnz"], GOLD["nnz_max"])
    assert a["rpt"].tolist() == GOLD["rpt"] and a["col"].tolist() == GOLD["col"]
    assert a["val"].tolist() == GOLD["val"] and a["val"].dtype == dtype


@pytest.mark.parametrize("name", sorted(CASES))
def test_reader_cases_match_reference_fixtures(name):
    case = CASES[name]
    with tempfile.NamedTemporaryFile("w", suffix=".mtx", delete=False) as f:
        f.write(case["text"])
    try:
        a = oracle.read_mtx(f.name, np.float64)
    finally:
        os.unlink(f.name)
    for k in ("M", "N", "nnz", "nnz_max"):
        assert a[k] == case[k], k
    assert a["rpt"].tolist() == case["rpt"] and a["col"].tolist() == case["col"] and a["val"].tolist() == case["val"]


def test_reader_against_reference_binary():
    """Live check against the reference's own nsparse.cu when oracle/_ref is present."""
    try:
        ref = oracle.ReferenceHost(np.float64)
    except (FileNotFoundError, OSError):
        pytest.skip("oracle/_ref not built")
    r = ref.read_mtx(MTX)
    a = oracle.read_mtx(MTX, np.float64)
    for k in ("rpt", "col", "val"):
        assert np.array_equal(r[k], a[k])


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_spgemm_golden(dtype):
    rpt, col = np.array(GOLD["rpt"], np.int32), np.array(GOLD["col"], np.int32)
    val = np.array(GOLD["val"], dtype)
    assert oracle.spgemm_flop(rpt, col, rpt) == GOLD["flop"] == 38
    assert oracle.spgemm_intprod(rpt, col, rpt).tolist() == GOLD["intprod_per_row"]
    for acc in (False, True):
        c_rpt, c_col, c_val = oracle.spgemm(rpt, col, val, rpt, col, val, acc_double=acc)
        assert c_rpt.tolist() == GOLD["c_rpt"] and c_col.tolist() == GOLD["c_col"]
        assert c_val.tolist() == GOLD["c_val"]          # exact in fp32 and fp64


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_spmv_golden_and_reference(dtype):
    rpt, col = np.array(GOLD["rpt"], np.int32), np.array(GOLD["col"], np.int32)
    val, x = np.array(GOLD["val"], dtype), np.array(GOLD["x"], dtype)
    y = oracle.spmv_csr(rpt, col, val, x)
    assert y.tolist() == GOLD["y"] == [13.0, 40.0, 101.0, 160.0, 256.0]
    assert np.array_equal(oracle.spmv_csr(rpt, col, val, x, parallel=True), y)
    try:
        ref = oracle.ReferenceHost(dtype)
    except (FileNotFoundError, OSError):
        return
    rng = np.random.default_rng(3)
    a = _rand_csr(300, 200, 0.05, 1, dtype)
    a.data = rng.random(a.nnz).astype(dtype)
    xx = rng.random(200).astype(dtype)
    # bit-exact: same left-to-right accumulation in `real`
    assert np.array_equal(ref.csr_kernel(a.indptr, a.indices, a.data, xx),
                          oracle.spmv_csr(a.indptr, a.indices, a.data, xx))


@pytest.mark.parametrize("m,k,n,dens,seed", [(50, 40, 60, 0.1, 0), (200, 200, 200, 0.05, 1), (1, 1, 1, 1.0, 2),
                                             (64, 300, 17, 0.2, 3), (500, 500, 500, 0.01, 4)])
def test_spgemm_vs_scipy(m, k, n, dens, seed):
    a, b = _rand_csr(m, k, dens, seed), _rand_csr(k, n, dens, seed + 100)
    c = (a @ b).tocsr()
    c.sort_indices()
    c_rpt, c_col, c_val = oracle.spgemm(a.indptr, a.indices, a.data, b.indptr, b.indices, b.data)
    # integer-valued inputs: no numerical cancellation to zero is possible (all positive), so
    # SciPy's structure equals the structural product
    assert np.array_equal(c_rpt, c.indptr) and np.array_equal(c_col, c.indices)
    assert np.array_equal(c_val, c.data)
    ok, msg = oracle.check_spgemm_answer((c_rpt, c_col, c_val), (c.indptr, c.indices, c.data))
    assert ok, msg


def test_spgemm_keeps_numerical_zeros_and_unsorted_input():
    # A = [1 -1], B = [[1],[1]]  ->  C = [0] structurally present (reference keeps it: no pruning)
    c_rpt, c_col, c_val = oracle.spgemm(np.array([0, 2]), np.array([0, 1]), np.array([1.0, -1.0]),
                                        np.array([0, 1, 2]), np.array([0, 0]), np.array([1.0, 1.0]))
    assert c_rpt.tolist() == [0, 1] and c_col.tolist() == [0] and c_val.tolist() == [0.0]
    # unsorted rows on input still give sorted output
    a = _rand_csr(80, 80, 0.1, 7)
    rng = np.random.default_rng(0)
    col, val = a.indices.copy(), a.data.copy()
    for i in range(80):
        s, e = a.indptr[i], a.indptr[i + 1]
        p = rng.permutation(e - s)
        col[s:e], val[s:e] = col[s:e][p], val[s:e][p]
    r1 = oracle.spgemm(a.indptr, col, val, a.indptr, col, val)
    r0 = oracle.spgemm(a.indptr, a.indices, a.data, a.indptr, a.indices, a.data)
    for x, y in zip(r0, r1):
        assert np.array_equal(x, y)


def test_spgemm_row_block_matches_full():
    a = _rand_csr(120, 90, 0.08, 11)
    b = _rand_csr(90, 70, 0.08, 12)
    full = oracle.spgemm(a.indptr, a.indices, a.data, b.indptr, b.indices, b.data)
    part = oracle.spgemm(a.indptr, a.indices, a.data, b.indptr, b.indices, b.data, rows=(30, 77))
    lo, hi = full[0][30], full[0][77]
    assert np.array_equal(part[0], full[0][30:78] - lo)
    assert np.array_equal(part[1], full[1][lo:hi]) and np.array_equal(part[2], full[2][lo:hi])


def test_comparators():
    r, c, v = np.array([0, 2]), np.array([0, 3]), np.array([1.0, 2.0])
    assert oracle.check_spgemm_answer((r, c, v), (r, c, v))[0]
    assert not oracle.check_spgemm_answer((r, c, v * (1 + 1e-9)), (r, c, v))[0]          # 1e-12 gate
    assert oracle.check_spgemm_answer((r, c, (v * (1 + 1e-9)).astype(np.float32)),
                                      (r, c, v.astype(np.float32)))[0]
    assert not oracle.check_spgemm_answer((r, np.array([0, 2]), v), (r, c, v))[0]
    assert not oracle.ans_check(np.array([1.0]), np.array([1.0 + 1e-9]))[0]
