"""GPU parity tests of the hash SpGEMM against the CPU oracle, through the C ABI.

Bar (BASELINE.json north_star): nnz, row pointer and column indices bit-exact; values within
1e-6 relative (fp32) / 1e-12 (fp64).  Integer-valued inputs make every sum exact, so for those the
values are required to be BIT-exact as well (any summation order gives the same float).
"""
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import oracle

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "test_mtx.json")))


@pytest.fixture(scope="module")
def ns():
    import nsparse_b200 as ns

    ns.load_library()
    return ns


@pytest.fixture(scope="module")
def ctx(ns):
    return ns.Context(0)


def _run(ns, ctx, a, b):
    a.memcpy()
    b.memcpy()
    c = ns.spgemm_kernel_hash(a, b, ctx)
    ctx.sync()
    return c, c.to_host()


def _oracle(a, b):
    return oracle.spgemm(a.rpt, a.col, a.val, b.rpt, b.col, b.val, acc_double=True)


def _check(ns, ctx, a, b, exact_values=False):
    c, got = _run(ns, ctx, a, b)
    want = _oracle(a, b)
    assert c.nnz == int(want[0][-1])
    assert c.intprod * 2 == oracle.spgemm_flop(a.rpt, a.col, b.rpt)
    ok, msg = oracle.check_spgemm_answer(got, want)
    assert ok, msg
    if exact_values:
        assert np.array_equal(got[2], want[2])
    return c, got


def _rand(ns, m, n, density, seed, dtype, ints=True, sort=True):
    rng = np.random.default_rng(seed)
    a = sp.random(m, n, density=density, random_state=rng, format="csr", dtype=np.float64)
    a.data = rng.integers(1, 4, size=a.nnz).astype(dtype) if ints else rng.random(a.nnz).astype(dtype)
    a.sort_indices()
    return ns.CSR.from_scipy(a, dtype)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_golden_test_mtx(ns, ctx, dtype):
    """Config C1: data/test.mtx, C = A^2, against the committed golden vectors."""
    a = ns.CSR(GOLD["M"], GOLD["N"], GOLD["rpt"], GOLD["col"], np.array(GOLD["val"], dtype))
    a.memcpy()
    assert ns.get_spgemm_flop(a, a, ctx) == GOLD["flop"]
    c, (rpt, col, val) = _run(ns, ctx, a, a)
    assert c.nnz == GOLD["c_nnz"] and c.intprod * 2 == GOLD["flop"]
    assert rpt.tolist() == GOLD["c_rpt"] and col.tolist() == GOLD["c_col"] and val.tolist() == GOLD["c_val"]


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("m,k,n,dens", [(1, 1, 1, 1.0), (37, 53, 41, 0.2), (300, 300, 300, 0.03),
                                        (2000, 1500, 1800, 0.01), (64, 4000, 64, 0.3)])
def test_random_exact(ns, ctx, dtype, m, k, n, dens):
    a, b = _rand(ns, m, k, dens, 1, dtype), _rand(ns, k, n, dens, 2, dtype)
    _check(ns, ctx, a, b, exact_values=True)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_random_float_values_tolerance(ns, ctx, dtype):
    a, b = _rand(ns, 3000, 3000, 0.002, 5, dtype, ints=False), _rand(ns, 3000, 3000, 0.002, 6, dtype, ints=False)
    _check(ns, ctx, a, b)


def test_empty_rows_and_empty_matrix(ns, ctx):
    z = ns.CSR(5, 5, np.zeros(6, np.int32), np.zeros(0, np.int32), np.zeros(0, np.float64))
    c, got = _run(ns, ctx, z, z)
    assert c.nnz == 0 and got[0].tolist() == [0] * 6
    # rows 1 and 3 of A empty; column 2 of A hits an empty row of B
    a = ns.CSR(4, 4, [0, 2, 2, 3, 3], [0, 2, 1], np.array([1.0, 2.0, 3.0]))
    b = ns.CSR(4, 3, [0, 1, 3, 3, 3], [2, 0, 1], np.array([5.0, 6.0, 7.0]))
    _check(ns, ctx, a, b, exact_values=True)
    m0 = ns.CSR(0, 7, np.zeros(1, np.int32), np.zeros(0, np.int32), np.zeros(0, np.float32))
    b7 = _rand(ns, 7, 9, 0.3, 3, np.float32)
    c, got = _run(ns, ctx, m0, b7)
    assert c.nnz == 0 and got[0].tolist() == [0]


def test_numerical_zero_is_kept(ns, ctx):
    a = ns.CSR(1, 2, [0, 2], [0, 1], np.array([1.0, -1.0]))
    b = ns.CSR(2, 1, [0, 1, 2], [0, 0], np.array([1.0, 1.0]))
    c, got = _run(ns, ctx, a, b)
    assert got[0].tolist() == [0, 1] and got[1].tolist() == [0] and got[2].tolist() == [0.0]


def test_unsorted_input_rows(ns, ctx):
    """The reference reader emits unsorted rows for symmetric files (nsparse.cu:115-123)."""
    a = _rand(ns, 400, 400, 0.03, 9, np.float64)
    rng = np.random.default_rng(1)
    col, val = a.col.copy(), a.val.copy()
    for i in range(a.M):
        s, e = a.rpt[i], a.rpt[i + 1]
        p = rng.permutation(e - s)
        col[s:e], val[s:e] = col[s:e][p], val[s:e][p]
    u = ns.CSR(a.M, a.N, a.rpt, col, val)
    _check(ns, ctx, u, u, exact_values=True)


def _row_with(nprod, n, seed):
    """A (1 x k) * B (k x n) whose single row has exactly `nprod` intermediate products."""
    import nsparse_b200 as ns

    rng = np.random.default_rng(seed)
    nb = 16 if (nprod % 16 == 0 and nprod >= 64) else 1
    la = nprod // nb
    k = max(la, 64)
    acol = np.sort(rng.choice(k, size=la, replace=False)).astype(np.int32)
    a = ns.CSR(1, k, [0, la], acol, np.ones(la, np.float64))
    bc = np.sort(rng.integers(0, n - nb, size=(k, nb)), axis=1) + np.arange(nb)   # distinct, sorted
    b = ns.CSR(k, n, np.arange(k + 1, dtype=np.int32) * nb, bc.reshape(-1).astype(np.int32), np.ones(k * nb))
    return a, b


@pytest.mark.parametrize("nprod", [32, 33, 64, 512, 513, 1024, 4096, 4097, 8192, 16384, 16385, 32768, 40000])
def test_symbolic_bin_boundaries(ns, ctx, nprod):
    """Rows whose intermediate-product count sits on every class edge of the symbolic ladder
    (reference edges 32/512/1024/2048/4096/8192, ours 32/512/4096/16384)."""
    a, b = _row_with(nprod, 200000, nprod)
    _check(ns, ctx, a, b, exact_values=True)


@pytest.mark.parametrize("nnz_row", [16, 17, 256, 257, 2048, 2049, 8192, 8193, 20000])
def test_numeric_bin_boundaries(ns, ctx, nnz_row):
    """Rows whose nnz(C_i) sits exactly on every class edge of the numeric ladder: B = identity-like
    so nnz(C_i) == nnz(A_i)."""
    n = 30000
    rng = np.random.default_rng(nnz_row)
    acol = np.sort(rng.choice(n, size=nnz_row, replace=False)).astype(np.int32)
    a = ns.CSR(2, n, [0, nnz_row, nnz_row + 1], np.concatenate([acol, [5]]).astype(np.int32),
               rng.integers(1, 5, nnz_row + 1).astype(np.float64))
    b = ns.CSR(n, n, np.arange(n + 1, dtype=np.int32), np.arange(n, dtype=np.int32), np.full(n, 2.0))
    c, got = _check(ns, ctx, a, b, exact_values=True)
    assert got[0].tolist() == [0, nnz_row, nnz_row + 1]


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_forced_bitmap_agrees_with_hash_classes(ns, dtype):
    """The bitmap kernels forced on for all rows above the 4-thread class must agree with the hash kernels
    and the oracle."""
    a = _rand(ns, 1500, 1200, 0.02, 21, dtype)
    b = _rand(ns, 1200, 5000, 0.01, 22, dtype)
    want = _oracle(a, b)
    for force in (False, True):
        c2 = ns.Context(0)
        if force:
            c2.set_option("sym_bitmap_min", 32)
            c2.set_option("num_bitmap_min", 16)
        _, got = _run(ns, c2, a, b)
        ok, msg = oracle.check_spgemm_answer(got, want)
        assert ok, f"force_bitmap={force}: {msg}"
        assert np.array_equal(got[2], want[2])
        c2.close()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_rmat_scale14(ns, ctx, dtype):
    """Small sibling of config C2 (heavy-tailed rows exercise every class incl. the bitmap kernels)."""
    from nsparse_b200 import gen

    a = gen.rmat_csr(14, 16, seed=12345, dtype=dtype, values="small_int")
    c, got = _check(ns, ctx, a, a, exact_values=True)
    assert c.nnz > 10 * a.nnz


def test_wide_matrix_multi_tile_bitmap(ns):
    """N larger than one bitmap tile (symbolic tile ~1.85 M columns, numeric ~1.23 M): heavy rows
    take several column passes."""
    n = 4_000_000
    rng = np.random.default_rng(4)
    la = 3000
    acol = np.sort(rng.choice(5000, size=la, replace=False)).astype(np.int32)
    a = ns.CSR(3, 5000, [0, la, la + 2, la + 2], np.concatenate([acol, [1, 7]]).astype(np.int32),
               rng.integers(1, 3, la + 2).astype(np.float64))
    bc = np.sort(rng.integers(0, n, size=(5000, 12)), axis=1)
    bc[:, 1:] = np.where(bc[:, 1:] <= bc[:, :-1], bc[:, :-1] + 1, bc[:, 1:])
    bc = np.minimum(bc, n - 1)
    keep = np.ones_like(bc, bool)
    keep[:, 1:] = bc[:, 1:] > bc[:, :-1]
    rpt = np.zeros(5001, np.int32)
    rpt[1:] = np.cumsum(keep.sum(axis=1))
    b = ns.CSR(5000, n, rpt, bc[keep].astype(np.int32), np.ones(int(rpt[-1])))
    c2 = ns.Context(0)
    _check(ns, c2, a, b, exact_values=True)
    c2.close()


def _heavy_case(ns, dtype, n=200_000, seed=7, unsorted=False):
    """A few rows of every kind the heavy (bitmap) class distinguishes: a long-B-row product with tens of
    thousands of outputs, a row with more than 1024 entries of A (several slabs), a light row."""
    rng = np.random.default_rng(seed)
    k = 4000
    blen = rng.integers(1, 60, size=k)
    blen[:40] = rng.integers(2000, 6000, size=40)          # hub rows of B
    rpt = np.zeros(k + 1, np.int64)
    rpt[1:] = np.cumsum(blen)
    bcol = np.concatenate([np.sort(rng.choice(n, size=int(l), replace=False)) for l in blen]).astype(np.int32)
    # a dense run so that consecutive columns share bitmap words
    bcol[rpt[3]:rpt[3] + 1500] = np.arange(1000, 2500)
    bcol[rpt[3]:rpt[4]] = np.unique(np.concatenate([bcol[rpt[3]:rpt[3] + 1500],
                                                    rng.choice(np.arange(3000, n), size=int(blen[3]) - 1500,
                                                               replace=False)]))[:blen[3]]
    bval = rng.integers(1, 4, size=int(rpt[-1])).astype(dtype)
    if unsorted:
        for i in range(k):
            p = rng.permutation(int(blen[i]))
            bcol[rpt[i]:rpt[i + 1]] = bcol[rpt[i]:rpt[i + 1]][p]
    b = ns.CSR(k, n, rpt.astype(np.int32), bcol, bval)
    rows = [np.sort(rng.choice(40, size=30, replace=False)),                 # 30 hub rows: ~100 k products
            np.sort(rng.choice(k, size=2500, replace=False)),                # 2500 entries: three slabs
            np.sort(rng.choice(np.arange(40, k), size=5, replace=False)),    # light
            np.array([3]),                                                    # the dense-run row alone
            np.sort(rng.choice(k, size=1024, replace=False))]                # exactly one slab
    arpt = np.zeros(len(rows) + 1, np.int64)
    arpt[1:] = np.cumsum([len(r) for r in rows])
    a = ns.CSR(len(rows), k, arpt.astype(np.int32), np.concatenate(rows).astype(np.int32),
               rng.integers(1, 3, size=int(arpt[-1])).astype(dtype))
    return a, b


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("opts", [
    {},                                                            # default windows / chunk size
    {"num_cap": 1024, "num_bitmap_min": 16, "sym_bitmap_min": 32},  # dozens of chunks per row
    {"num_window_shift": 16, "sym_window_shift": 16, "num_bitmap_min": 16, "sym_bitmap_min": 32},   # 4 windows
    {"num_window_shift": 16, "sym_window_shift": 16, "num_cap": 256, "no_vec": 1, "num_bitmap_min": 16,
     "sym_bitmap_min": 32},                                        # windows x chunks, scalar mark loads
])
def test_heavy_class_windows_chunks_slabs(ns, dtype, opts):
    """Every path of the bitmap kernels: column windows, accumulator chunks (cursor search), rows with more
    than 1024 A entries (red.global mode), dense runs, the scalar fallback of the 128-bit mark loads."""
    a, b = _heavy_case(ns, dtype)
    c2 = ns.Context(0)
    for k, v in opts.items():
        c2.set_option(k, v)
    _check(ns, c2, a, b, exact_values=True)
    c2.close()


@pytest.mark.parametrize("opts", [{}, {"num_cap": 512, "num_window_shift": 16, "sym_window_shift": 16,
                                       "num_bitmap_min": 16, "sym_bitmap_min": 32}])
def test_heavy_class_unsorted_b(ns, opts):
    """Rows of B not column-sorted (reader of symmetric files, nsparse.cu:115-123): the kernels detect it
    and filter by column range instead of searching."""
    a, b = _heavy_case(ns, np.float64, unsorted=True)
    c2 = ns.Context(0)
    for k, v in opts.items():
        c2.set_option(k, v)
    _check(ns, c2, a, b, exact_values=True)
    c2.close()


def test_rpt32_narrowing_and_overflow_guard(ns, ctx):
    import ctypes as C

    import torch

    a = _rand(ns, 100, 100, 0.1, 1, np.float32)
    c, got = _run(ns, ctx, a, a)
    r32 = torch.empty(a.M + 1, dtype=torch.int32, device="cuda")
    ctx.check(ctx.lib.nsp_rpt64_to_rpt32(ctx.handle, a.M, C.c_void_p(c.d_rpt64.data_ptr()), c.nnz,
                                         C.c_void_p(r32.data_ptr())))
    ctx.sync()
    assert np.array_equal(r32.cpu().numpy().astype(np.int64), got[0])
    rc = ctx.lib.nsp_rpt64_to_rpt32(ctx.handle, a.M, C.c_void_p(c.d_rpt64.data_ptr()), 2 ** 31,
                                    C.c_void_p(r32.data_ptr()))
    assert rc == -3   # NSP_ERR_OVERFLOW instead of the reference's silent wrap


def test_host_buffer_entry_point(ns, ctx):
    """nsp_spgemm_host_*: host CSR in, device result fetched into host arrays."""
    import ctypes as C

    a = _rand(ns, 500, 400, 0.03, 31, np.float64)
    b = _rand(ns, 400, 600, 0.03, 32, np.float64)
    p = lambda x: x.ctypes.data_as(C.c_void_p)
    nnz = C.c_longlong()
    ctx.check(ctx.lib.nsp_spgemm_host_d(ctx.handle, a.M, a.N, b.N, p(a.rpt), p(a.col), p(a.val), p(b.rpt),
                                        p(b.col), p(b.val), C.byref(nnz)))
    want = _oracle(a, b)
    assert nnz.value == int(want[0][-1])
    rpt = np.empty(a.M + 1, np.int64)
    col = np.empty(nnz.value, np.int32)
    val = np.empty(nnz.value, np.float64)
    ctx.check(ctx.lib.nsp_spgemm_host_fetch_d(ctx.handle, p(rpt), p(col), p(val)))
    ok, msg = oracle.check_spgemm_answer((rpt, col, val), want)
    assert ok, msg
    ctx.check(ctx.lib.nsp_spgemm_host_release(ctx.handle))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_host_stream_entry_point(ns, ctx, dtype):
    """nsp_spgemm_host_stream_*: host CSR in, C streamed to the host while it is computed (row ranges + a
    drain thread).  The fold of the drained chunks must equal the fold of the oracle's C cut the same way."""
    import ctypes as C

    import torch

    a = _rand(ns, 900, 700, 0.03, 41, dtype)
    b = _rand(ns, 700, 1100, 0.03, 42, dtype)
    want = _oracle(a, b)
    w_rpt, w_col, w_val = want[0].astype(np.int64), want[1].astype(np.int32), want[2].astype(dtype)
    p = lambda x: x.ctypes.data_as(C.c_void_p)
    stage = torch.empty(8192, dtype=torch.uint8).pin_memory()
    half, pieces = 4096, 3
    nnz, nbytes, csum = C.c_longlong(), C.c_longlong(), C.c_ulonglong()
    fn = ctx.lib.nsp_spgemm_host_stream_d if dtype == np.float64 else ctx.lib.nsp_spgemm_host_stream_s
    ctx.check(fn(ctx.handle, a.M, a.N, b.N, p(a.rpt), p(a.col), p(a.val), p(b.rpt), p(b.col), p(b.val),
                 C.c_void_p(stage.data_ptr()), stage.numel(), pieces, C.byref(nnz), C.byref(csum), C.byref(nbytes)))
    assert nnz.value == int(w_rpt[-1])
    total = int(w_rpt[-1])
    rows = [int(np.searchsorted(w_rpt, total // pieces * k, side="left")) for k in range(pieces)] + [a.M]
    ranges = [w_rpt.tobytes()]
    for k in range(pieces):
        e0, e1 = int(w_rpt[rows[k]]), int(w_rpt[rows[k + 1]])
        ranges += [w_col[e0:e1].tobytes(), w_val[e0:e1].tobytes()]
    fold, moved = 0, 0
    for r in ranges:
        for off in range(0, len(r), half):
            chunk = r[off:off + half]
            x = int.from_bytes(chunk[:8].ljust(8, b"\0"), "little")
            y = int.from_bytes(chunk[-8:], "little") if len(chunk) >= 8 else 0
            fold = (fold * 1099511628211 + (x ^ ((y << 1) & 0xFFFFFFFFFFFFFFFF))) & 0xFFFFFFFFFFFFFFFF
            moved += len(chunk)
    assert nbytes.value == moved
    assert csum.value == fold
    ctx.check(ctx.lib.nsp_spgemm_host_release(ctx.handle))


def test_host_drain_fold(ns, ctx):
    """nsp_spgemm_host_s + nsp_spgemm_host_drain (the e2e leg of bench.py): the fold of the staged chunks of
    rpt, col, val must equal the fold of the oracle's arrays."""
    import ctypes as C

    import torch

    a = _rand(ns, 600, 500, 0.04, 51, np.float32)
    b = _rand(ns, 500, 800, 0.04, 52, np.float32)
    want = _oracle(a, b)
    p = lambda x: x.ctypes.data_as(C.c_void_p)
    nnz, nbytes, csum = C.c_longlong(), C.c_longlong(), C.c_ulonglong()
    ctx.check(ctx.lib.nsp_spgemm_host_s(ctx.handle, a.M, a.N, b.N, p(a.rpt), p(a.col), p(a.val), p(b.rpt), p(b.col),
                                        p(b.val), C.byref(nnz)))
    stage = torch.empty(8192, dtype=torch.uint8).pin_memory()
    ctx.check(ctx.lib.nsp_spgemm_host_drain(ctx.handle, C.c_void_p(stage.data_ptr()), stage.numel(), C.byref(csum),
                                            C.byref(nbytes)))
    fold, moved = 0, 0
    for r in (want[0].astype(np.int64).tobytes(), want[1].astype(np.int32).tobytes(), want[2].astype(np.float32).tobytes()):
        for off in range(0, len(r), 4096):
            chunk = r[off:off + 4096]
            x = int.from_bytes(chunk[:8].ljust(8, b"\0"), "little")
            y = int.from_bytes(chunk[-8:], "little") if len(chunk) >= 8 else 0
            fold = (fold * 1099511628211 + (x ^ ((y << 1) & 0xFFFFFFFFFFFFFFFF))) & 0xFFFFFFFFFFFFFFFF
            moved += len(chunk)
    assert nnz.value == int(want[0][-1]) and nbytes.value == moved and csum.value == fold
    ctx.check(ctx.lib.nsp_spgemm_host_release(ctx.handle))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("offset", [0, 12345])
@pytest.mark.parametrize("mode", ["dma", "dma_small_tiles", "dma_small_tiles_3peers", "dma_small_tiles_ranges", "ce_only", "sm_only",
                                  "sm_only_small_tiles", "sm_only_3peers", "tma"])
def test_tile_pusher_on_one_gpu(ns, dtype, offset, mode):
    """The multi-GPU allgatherv path (tile counters in every numeric kernel, then either the copy engines driven by
    the polling host thread plus SM stores for what is left when the kernels end, csrc/peer_dma.cu -- each of the two
    alone as well -- or the TMA pusher kernel, csrc/peer_push.cu) with the 'peer' being a second buffer on the SAME
    GPU: after the product the peer copy must equal C entry for entry at the block's displacement, and nothing
    outside the block may have been touched."""
    import ctypes as C

    import torch

    from nsparse_b200 import gen

    ctx = ns.Context(0)
    if mode == "tma":
        ctx.set_option("gather_tma", 1)
    if "small_tiles" in mode:
        ctx.set_option("dma_tile_log", 12)
    if mode == "ce_only":
        ctx.set_option("gather_sm", 0)
    if mode.endswith("ranges"):
        ctx.set_option("no_ranges", -1)      # the heavy rows through num_hash_ranges_kernel (it counts tiles too)
    if mode.startswith("sm_only"):
        ctx.set_option("gather_sm", 2)
    a = gen.rmat_csr(13, 16, seed=4, dtype=dtype, values="small_int")
    a.memcpy()
    d_rpt64, nnz, _ = ns.spgemm_symbolic(a, a, ctx)
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    total = offset + nnz + 999
    col = torch.full((total,), -7, dtype=torch.int32, device="cuda")
    val = torch.full((total,), -7, dtype=tdt, device="cuda")
    npeers = 3 if mode.endswith("3peers") else 1
    pcols = [torch.full((total,), -7, dtype=torch.int32, device="cuda") for _ in range(npeers)]
    pvals = [torch.full((total,), -7, dtype=tdt, device="cuda") for _ in range(npeers)]
    ctx.check(ctx.lib.nsp_spgemm_set_peers(ctx.handle, npeers, (C.c_void_p * npeers)(*[t.data_ptr() for t in pcols]),
                                           (C.c_void_p * npeers)(*[t.data_ptr() for t in pvals]), offset))
    try:
        ns.spgemm_numeric(a, a, d_rpt64, nnz, ctx, out=(col[offset:], val[offset:]))
    finally:
        ctx.check(ctx.lib.nsp_spgemm_set_peers(ctx.handle, 0, None, None, 0))
    err = C.c_int(0)
    ctx.check(ctx.lib.nsp_spgemm_peers_status(ctx.handle, C.byref(err)))
    assert err.value == 0
    want = oracle.spgemm(a.rpt, a.col, a.val, a.rpt, a.col, a.val, acc_double=True)
    assert np.array_equal(col[offset:offset + nnz].cpu().numpy(), want[1])
    assert np.array_equal(val[offset:offset + nnz].cpu().numpy(), want[2])
    for pcol, pval in zip(pcols, pvals):
        assert torch.equal(pcol, col) and torch.equal(pval, val)
        assert int((pcol[:offset] != -7).sum()) == 0 and int((pcol[offset + nnz:] != -7).sum()) == 0
    if mode != "tma":
        n_ce, n_sm = C.c_longlong(-1), C.c_longlong(-1)
        ctx.check(ctx.lib.nsp_spgemm_peers_stats(ctx.handle, C.byref(n_ce), C.byref(n_sm), None))
        assert n_ce.value >= 0 and n_sm.value >= 0 and n_ce.value + n_sm.value > 0
        if mode == "ce_only":
            assert n_sm.value == 0
        if mode.startswith("sm_only"):
            assert n_ce.value == 0
    ctx.close()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("opts", [{}, {"sym_window_shift": 16, "num_window_shift": 16, "num_cap": 256},
                                  {"no_seg": 1, "sym_window_shift": 16, "num_window_shift": 16, "num_cap": 256}])
def test_flat_traversal_of_the_heavy_kernels(ns, dtype, opts):
    """The traversal the bitmap kernels use when the B rows of the class are short (run_flat, configs C4 / C5), forced
    on (no_flat = -1) for inputs that take the part traversal by default: R-MAT A^2 with low class thresholds, one
    window and several windows x chunks, with and without the precomputed segments; and a product with really short
    B rows (4 per row) and a wide C where it is the default.  Bit-exact against the oracle."""
    from nsparse_b200 import gen

    ctx = ns.Context(0)
    ctx.set_option("no_flat", -1)
    ctx.set_option("sym_bitmap_min", 64)
    ctx.set_option("num_bitmap_min", 64)
    for k, v in opts.items():
        ctx.set_option(k, v)
    a = gen.rmat_csr(13, 16, seed=6, dtype=dtype, values="small_int")
    a.memcpy()
    c = ns.spgemm_kernel_hash(a, a, ctx)
    ctx.sync()
    want = oracle.spgemm(a.rpt, a.col, a.val, a.rpt, a.col, a.val, acc_double=True)
    got = c.to_host()
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2])
    ctx.close()
    # default choice on short B rows: long rows of A (up to 3000 entries) times 4 entries per row of B, N = 300000
    ctx = ns.Context(0)
    rng = np.random.default_rng(8)
    lens = np.r_[rng.integers(1, 60, 400), [1500, 3000, 2600, 1024, 1025]]
    K, N = 20000, 300000
    rpt = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    col = np.concatenate([np.sort(rng.choice(K, n, replace=False)) for n in lens]).astype(np.int32)
    a2 = ns.CSR(len(lens), K, rpt, col, rng.integers(1, 4, len(col)).astype(dtype))
    b2 = gen.er_csr(K, N, 4, seed=5, dtype=dtype, values="ones")
    a2.memcpy()
    b2.memcpy()
    c = ns.spgemm_kernel_hash(a2, b2, ctx)
    ctx.sync()
    want = oracle.spgemm(a2.rpt, a2.col, a2.val, b2.rpt, b2.col, b2.val, acc_double=True, n_cols=N)
    got = c.to_host()
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2])
    ctx.close()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("hash_order", [0, 1, 2])
def test_hash_row_ordering_paths(ns, dtype, hash_order):
    """The three ways a hash-class row is brought into column order -- buckets ordered inside shared memory (evenly
    spread columns, the default there), buckets ranked through C (hash_order = 2), bitonic sort of the table
    (hash_order = 1) -- on a product with evenly spread columns over a WIDE C (2^22 columns: every class of the hash
    ladder, tables of 32 .. 16384 slots) and on R-MAT A^2 (clustered columns).  Bit-exact against the oracle."""
    from nsparse_b200 import gen

    ctx = ns.Context(0)
    ctx.set_option("hash_order", hash_order)
    n = 1 << 22
    a = gen.powerlaw_csr(3000, mean_nnz=48, max_row=2000, seed=5, dtype=dtype, values="ones")
    a = type(a)(a.M, n, a.rpt, (a.col.astype(np.int64) * (n // 3000)).astype(np.int32), a.val, "wide_rows")
    b = gen.er_csr(n, n, 4, seed=6, dtype=dtype, values="ones")
    a.memcpy()
    b.memcpy()
    c = ns.spgemm_kernel_hash(a, b, ctx)
    ctx.sync()
    got = c.to_host()
    want = oracle.spgemm(a.rpt, a.col, a.val, b.rpt, b.col, b.val, acc_double=True, n_cols=n)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2])
    assert int(np.diff(want[0]).max()) > 4096, "no row reached the 1024-thread class"
    r = gen.rmat_csr(12, 16, seed=8, dtype=dtype, values="small_int")
    r.memcpy()
    c = ns.spgemm_kernel_hash(r, r, ctx)
    ctx.sync()
    got = c.to_host()
    want = oracle.spgemm(r.rpt, r.col, r.val, r.rpt, r.col, r.val, acc_double=True)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2])
    ctx.close()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_wide_product_hash_ranges(ns, dtype):
    """num_hash_ranges_kernel: rows above the hash ladder (more than 8192 entries) of a product with a WIDE C (2^22
    columns: 8 bitmap windows) and 4-entry B rows take hash passes over column ranges instead of the bitmap windows --
    rows of one range, of several ranges, and (forced with no_ranges = -1) R-MAT A^2, whose columns cluster so that
    ranges are single bins and the bitonic fallback orders them.  Bit-exact against the oracle, and against the bitmap
    kernels (no_ranges = 1) on the same input."""
    from nsparse_b200 import gen

    n = 1 << 22
    a = gen.powerlaw_csr(8000, mean_nnz=24, max_row=7000, seed=11, dtype=dtype, values="ones")
    # (rows of up to 7000 entries times 12-entry B rows: up to 84000 outputs; rows of more than 8 windows x 12288
    # entries stay with the bitmap kernels, the others take one to six ranges)
    a = type(a)(a.M, n, a.rpt, (a.col.astype(np.int64) * (n // 8000)).astype(np.int32), a.val, "wide_rows")
    b = gen.er_csr(n, n, 12, seed=12, dtype=dtype, values="ones")
    a.memcpy()
    b.memcpy()
    want = oracle.spgemm(a.rpt, a.col, a.val, b.rpt, b.col, b.val, acc_double=True, n_cols=n)
    per_row = np.diff(want[0])
    assert int(per_row.max()) > 65536 and int(((per_row > 2 * 12288) & (per_row <= 65536)).sum()) > 0
    assert int(((per_row > 8192) & (per_row <= 12288)).sum()) > 0
    for no_ranges in (0, 1):
        ctx = ns.Context(0)
        ctx.set_option("no_ranges", no_ranges)
        c = ns.spgemm_kernel_hash(a, b, ctx)
        ctx.sync()
        got = c.to_host()
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2])
        ctx.close()
    r = gen.rmat_csr(13, 16, seed=3, dtype=dtype, values="small_int")
    r.memcpy()
    want = oracle.spgemm(r.rpt, r.col, r.val, r.rpt, r.col, r.val, acc_double=True)
    ctx = ns.Context(0)
    ctx.set_option("no_ranges", -1)
    c = ns.spgemm_kernel_hash(r, r, ctx)
    ctx.sync()
    got = c.to_host()
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2])
    ctx.close()
