"""GPU parity tests of the AMB path through the C ABI: the device conversion must produce the oracle's
arrays bit for bit, and y = A x must match the CSR SpMV oracle within 1e-6 (fp32) / 1e-12 (fp64)."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import amb as A
from oracle import oracle

pytestmark = pytest.mark.gpu

ARRAYS = ("cs", "cl", "sellcs_col", "sellcs_val", "s_write_permutation", "s_write_permutation_offset",
          "write_permutation")
SCALARS = ("M", "N", "pad_M", "chunk", "SIGMA", "seg_size", "seg_num", "block_size", "c_size", "nnz")


@pytest.fixture(scope="module")
def ns():
    import nsparse_b200 as ns

    ns.load_library()
    return ns


@pytest.fixture(scope="module")
def ctx(ns):
    return ns.Context(0)


def _rand(ns, m, n, dens, seed, dtype, ints=False):
    rng = np.random.default_rng(seed)
    a = sp.random(m, n, density=dens, random_state=rng, format="csr", dtype=np.float64)
    a.data = rng.integers(1, 9, a.nnz).astype(dtype) if ints else (rng.random(a.nnz) + 0.5).astype(dtype)
    a.sort_indices()
    return ns.CSR.from_scipy(a, dtype)


def _tol(dtype):
    return 1e-6 if np.dtype(dtype) == np.float32 else 1e-12


def _convert_and_check(ns, ctx, a, seg, bs, x=None):
    import torch

    a.memcpy()
    plan = ns.Plan().set_plan(seg, bs) if seg else ns.Plan()
    amb = ns.csr2amb(a, plan, ctx=ctx)
    got = amb.to_host()
    want = A.convert_amb(a.rpt, a.col, a.val, a.M, a.N, got["seg_size"], got["block_size"])
    for k in SCALARS:
        assert got[k] == want[k], (k, got[k], want[k])
    for k in ARRAYS:
        assert np.array_equal(got[k], want[k]), k
    rng = np.random.default_rng(7)
    x = rng.random(a.N).astype(a.dtype) if x is None else x
    y = ns.spmv_amb(amb, torch.from_numpy(x).cuda(), ctx=ctx).cpu().numpy()
    y0 = oracle.spmv_csr(a.rpt, a.col, a.val.astype(np.float64), x.astype(np.float64))
    scale = np.abs(a.to_scipy()).astype(np.float64) @ np.abs(x.astype(np.float64))
    assert (np.abs(y.astype(np.float64) - y0) <= _tol(a.dtype) * np.maximum(scale, 1e-300)).all()
    return amb, plan


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_golden_test_mtx(ns, ctx, dtype):
    import json

    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "test_mtx.json")))
    m = sp.csr_matrix((np.array(g["val"], dtype), np.array(g["col"]), np.array(g["rpt"])), shape=(5, 5))
    m.sort_indices()
    a = ns.CSR.from_scipy(m, dtype)
    import torch

    for seg, bs in [(65536, 1), (1, 1), (2, 3), (4, 20), (0, 0)]:
        a.memcpy()
        plan = ns.Plan().set_plan(seg, bs) if seg else ns.Plan()
        amb = ns.csr2amb(a, plan, ctx=ctx)
        tdt = torch.float32 if dtype == np.float32 else torch.float64
        y = ns.spmv_amb(amb, torch.tensor(g["x"], dtype=tdt).cuda(), ctx=ctx)
        assert y.cpu().tolist() == g["y"]


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("m,n,dens", [(1, 1, 1.0), (33, 70, 0.2), (1000, 3000, 0.01), (40000, 300, 0.02),
                                      (70000, 150000, 0.0002)])
@pytest.mark.parametrize("seg,bs", [(65536, 1), (1024, 2), (4096, 5), (2048, 20)])
def test_conversion_matches_oracle(ns, ctx, dtype, m, n, dens, seg, bs):
    _convert_and_check(ns, ctx, _rand(ns, m, n, dens, m + n, dtype), seg, bs)


@pytest.mark.parametrize("bs", range(1, 21))
def test_every_block_size(ns, ctx, bs):
    _convert_and_check(ns, ctx, _rand(ns, 3000, 2500, 0.01, bs, np.float64), 1024, bs)


def test_planner_matches_reference_footprint_model(ns, ctx):
    from nsparse_b200 import gen

    lap = gen.laplacian5_csr(48)
    amb, plan = _convert_and_check(ns, ctx, lap, 0, 0)
    seg, bs, _ = A.plan_footprint(lap.rpt, lap.col, lap.val, lap.M, lap.N)
    assert (plan.seg_size, plan.block_size) == (seg, bs) and plan.isPlan
    a = _rand(ns, 2000, 5000, 0.01, 3, np.float32)
    amb, plan = _convert_and_check(ns, ctx, a, 0, 0)
    seg, bs, _ = A.plan_footprint(a.rpt, a.col, a.val, a.M, a.N)
    assert (plan.seg_size, plan.block_size) == (seg, bs)


@pytest.mark.parametrize("seg,bs", [(1024, 1), (1024, 4), (65536, 7)])
def test_unsorted_rows_take_the_sorted_copy_path(ns, ctx, seg, bs):
    """Rows whose columns are in arbitrary order (what the reference reader can emit): the entries of
    a virtual row are sorted by column first, so y is exact (the reference loses entries here)."""
    a = _rand(ns, 500, 3000, 0.02, 11, np.float64)
    rng = np.random.default_rng(2)
    col, val = a.col.copy(), a.val.copy()
    for i in range(a.M):
        s, e = a.rpt[i], a.rpt[i + 1]
        p = rng.permutation(e - s)
        col[s:e], val[s:e] = col[s:e][p], val[s:e][p]
    u = ns.CSR(a.M, a.N, a.rpt, col, val)
    _convert_and_check(ns, ctx, u, seg, bs)


@pytest.mark.parametrize("bs", [1, 3])
def test_duplicate_entries_are_summed(ns, ctx, bs):
    """The reference reader does not merge duplicates; y = A x must add both."""
    a = _rand(ns, 300, 2000, 0.02, 13, np.float64, ints=True)
    rpt = np.zeros(a.M + 1, np.int64)
    cols, vals = [], []
    for i in range(a.M):
        s, e = a.rpt[i], a.rpt[i + 1]
        c, v = list(a.col[s:e]), list(a.val[s:e])
        if len(c) >= 2:                      # repeat the first two columns of the row
            c += c[:2]
            v += [1.0, 2.0]
        cols += c
        vals += v
        rpt[i + 1] = len(cols)
    d = ns.CSR(a.M, a.N, rpt.astype(np.int32), np.array(cols, np.int32), np.array(vals))
    _convert_and_check(ns, ctx, d, 1024, bs)


def test_empty_rows_empty_matrix_and_m_above_65536(ns, ctx):
    import torch

    z = ns.CSR(5, 5, np.zeros(6, np.int32), np.zeros(0, np.int32), np.zeros(0, np.float64))
    z.memcpy()
    amb = ns.csr2amb(z, ctx=ctx)
    assert amb.c_size == 0
    assert ns.spmv_amb(amb, torch.ones(5, dtype=torch.float64, device="cuda"), ctx=ctx).cpu().tolist() == [0.0] * 5
    # M > 65536 exercises the high part of the write permutation; M % 32 != 0 the padding rows
    a = _rand(ns, 140003, 1000, 0.003, 5, np.float64)
    _convert_and_check(ns, ctx, a, 65536, 1)
    _convert_and_check(ns, ctx, a, 1024, 3)


def test_blocks_touching_last_column_do_not_read_past_x(ns, ctx):
    """Entries in column N-1 with block_size 20: x has exactly N entries and is followed by NaNs."""
    import torch

    n = 257
    a = sp.csr_matrix((np.ones(64), (np.arange(64), np.full(64, n - 1))), shape=(64, n))
    a = (a + sp.eye(64, n, format="csr")).tocsr()
    a.sort_indices()
    c = ns.CSR.from_scipy(a, np.float64)
    c.memcpy()
    amb = ns.csr2amb(c, ns.Plan().set_plan(65536, 20), ctx=ctx)
    buf = torch.full((n + 64,), float("nan"), dtype=torch.float64, device="cuda")
    buf[:n] = 1.0
    y = ns.spmv_amb(amb, buf[:n], ctx=ctx).cpu().numpy()
    assert np.array_equal(y, np.asarray(a.sum(axis=1)).ravel())


def test_laplacian_mid_size(ns, ctx):
    """Small sibling of config C3 (5-point Laplacian): many segments, M a multiple of the window."""
    from nsparse_b200 import gen

    lap = gen.laplacian5_csr(512)          # M = N = 262144 -> seg_size 65536 forced, 4 segments
    amb, plan = _convert_and_check(ns, ctx, lap, 0, 0)
    assert plan.seg_size == 65536 and plan.block_size == 1


def test_autotune_path_and_host_entry_point(ns, ctx):
    import ctypes as C

    import torch

    a = _rand(ns, 3000, 4000, 0.01, 9, np.float32)
    a.memcpy()
    x = np.random.default_rng(1).random(a.N).astype(np.float32)
    amb = ns.csr2amb(a, ns.Plan(), x=torch.from_numpy(x).cuda(), autotune=True, ctx=ctx)
    assert 1 <= amb.block_size <= 20 and amb.seg_size in (65536, 1024, 2048, 3072, 4096)
    y = np.empty(a.M, np.float32)
    ctx.check(ctx.lib.nsp_spmv_amb_host_s(ctx.handle, C.byref(amb._c), x.ctypes.data_as(C.c_void_p),
                                          y.ctypes.data_as(C.c_void_p)))
    y0 = oracle.spmv_csr(a.rpt, a.col, a.val.astype(np.float64), x.astype(np.float64))
    assert np.allclose(y, y0, rtol=1e-5)
