"""SpGEMM parity PINNED TO THE REFERENCE: tests/golden/spgemm_ref_*.npz hold what the reference's own GPU
SpGEMM (cuda-cpp/inc/HashSpGEMM_volta.hpp:974-1010 SpGEMM_Hash, header unmodified, built by
`make -C oracle ref_spgemm`) computed on a B200 (tests/golden/make_spgemm_golden.py ran it under gpurun;
the cuda-c tree livelocks on sm_100 beyond the pwarp class and agrees on data/test.mtx, see
spgemm_ref_manifest.json).

  * not gpu: the CPU oracle (oracle/oracle.c) reproduces every fixture -- nnz, rpt, col exact, values within
    1e-6 / 1e-12 relative, BIT-exact where the inputs make every sum exact;
  * gpu: the CUDA product through the C ABI reproduces the same fixtures.
"""
import glob
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "spgemm_ref_*.npz")))
IDS = [os.path.basename(f)[len("spgemm_ref_"):-4] for f in FIXTURES]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load_inputs(fx):
    """(a, b) as dicts M, N, rpt, col, val -- from the fixture, from test.mtx, or regenerated from the seed."""
    kind = str(fx["kind"])
    dt = fx["c_val"].dtype if "c_val" in fx.files else fx["sample_val"].dtype
    if kind == "mtx":
        m = oracle.read_mtx(os.path.join(GOLDEN, "test.mtx"), dt)
        return m, m
    if kind == "full":
        a = dict(M=int(fx["a_M"]), N=int(fx["a_N"]), rpt=fx["a_rpt"], col=fx["a_col"], val=fx["a_val"])
        b = dict(M=int(fx["b_M"]), N=int(fx["b_N"]), rpt=fx["b_rpt"], col=fx["b_col"], val=fx["b_val"]) if "b_rpt" in fx.files else a
        return a, b
    from nsparse_b200 import gen

    g = json.loads(str(fx["gen"]))
    m = gen.rmat_csr(g["scale"], g["ef"], seed=g["seed"], dtype=dt, values=g["values"])
    assert sha(m.rpt) + sha(m.col) + sha(m.val) == str(fx["a_sha"]), "the generator no longer reproduces the fixture's input"
    a = dict(M=m.M, N=m.N, rpt=m.rpt, col=m.col, val=m.val)
    return a, a


def compare(fx, got):
    """got = (rpt int64, col, val) of the whole product."""
    rpt, col, val = got
    tol = oracle.TOL[np.dtype(val.dtype)]
    if str(fx["kind"]) in ("mtx", "full"):
        ok, msg = oracle.check_spgemm_answer(got, (fx["c_rpt"], fx["c_col"], fx["c_val"]))
        assert ok, msg
        return
    assert int(rpt[-1]) == int(fx["c_nnz"])
    assert np.array_equal(np.asarray(rpt, np.int64), fx["c_rpt"].astype(np.int64))
    assert sha(np.asarray(col, np.int32)) == str(fx["c_col_sha"])
    rows, srpt = fx["sample_rows"], fx["sample_rpt"]
    lens = np.diff(srpt)
    idx = np.repeat(np.asarray(rpt, np.int64)[rows] - srpt[:-1], lens) + np.arange(int(srpt[-1]))
    assert np.array_equal(col[idx], fx["sample_col"])
    g = json.loads(str(fx["gen"]))
    if g["values"] == "small_int":              # every sum is an exact small integer: bit-exact
        assert np.array_equal(val[idx], fx["sample_val"])
        rowsum = np.add.reduceat(np.r_[val.astype(np.float64), 0.0], np.minimum(np.asarray(rpt[:-1], np.int64), len(val)))
        rowsum[np.diff(rpt) == 0] = 0.0
        assert np.array_equal(rowsum, fx["rowsum"])
    else:
        d = np.abs(val[idx].astype(np.float64) - fx["sample_val"])
        assert (d <= tol * np.abs(fx["sample_val"].astype(np.float64))).all()


def test_fixtures_present():
    """The reference-generated goldens are committed (at least data/test.mtx and the R-MAT cases)."""
    assert any("test_mtx" in i for i in IDS) and any("rmat" in i for i in IDS), IDS
    m = json.load(open(os.path.join(GOLDEN, "spgemm_ref_manifest.json")))
    # the two reference trees agree where both run
    assert m["test_mtx_d"]["c_vs_cpp"] == "agree" and m["test_mtx_s"]["c_vs_cpp"] == "agree"


@pytest.mark.parametrize("path", FIXTURES, ids=IDS)
def test_oracle_matches_reference_gpu(path):
    fx = np.load(path)
    a, b = load_inputs(fx)
    got = oracle.spgemm(a["rpt"], a["col"], a["val"], b["rpt"], b["col"], b["val"], acc_double=True, n_cols=b["N"])
    compare(fx, got)


@pytest.mark.gpu
@pytest.mark.parametrize("path", FIXTURES, ids=IDS)
def test_cuda_matches_reference_gpu(path):
    import nsparse_b200 as ns

    fx = np.load(path)
    a, b = load_inputs(fx)
    A = ns.CSR(a["M"], a["N"], a["rpt"], a["col"], a["val"])
    B = A if b is a else ns.CSR(b["M"], b["N"], b["rpt"], b["col"], b["val"])
    A.memcpy()
    if B is not A:
        B.memcpy()
    ctx = ns.default_context(0)
    c = ns.spgemm_kernel_hash(A, B, ctx)
    ctx.sync()
    compare(fx, c.to_host())
