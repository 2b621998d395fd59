"""Synthetic CSR inputs of the BASELINE.json configs.

All generators are deterministic functions of their seed.  R-MAT edges come from the native
counter-based generator (csrc/gen.cpp, OpenMP) with ``rmat_edges_numpy`` as its bit-identical
numpy mirror (used by the tests to pin it); everything else is numpy.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .csr import CSR

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
    x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return x ^ (x >> np.uint64(31))


def rmat_edges_numpy(scale: int, n_edges: int, seed: int):
    """Mirror of nsp_gen_rmat_edges: Graph500 Kronecker, (a,b,c,d)=(.57,.19,.19,.05)."""
    with np.errstate(over="ignore"):
        key = _splitmix64(np.array([seed], dtype=np.uint64))[0]
        e = np.arange(n_edges, dtype=np.uint64)
        src = np.zeros(n_edges, dtype=np.uint64)
        dst = np.zeros(n_edges, dtype=np.uint64)
        ta, tab, tabc = 2448131358, 3264175144, 4080218931
        for l in range(scale):
            r = _splitmix64(key ^ (e * np.uint64(64) + np.uint64(l))) >> np.uint64(32)
            sbit = (r >= tab).astype(np.uint64)
            dbit = (((r >= ta) & (r < tab)) | (r >= tabc)).astype(np.uint64)
            src = (src << np.uint64(1)) | sbit
            dst = (dst << np.uint64(1)) | dbit
    return src.astype(np.int64), dst.astype(np.int64)


_GEN_LIB = None


def _gen_lib():
    """libnsparse_gen.so: the generator alone (csrc/gen.cpp, no CUDA), so that generating an input -- in
    bench.py's reference arm, in the CPU tests -- does not load the product library."""
    global _GEN_LIB
    if _GEN_LIB is None:
        import os

        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libnsparse_gen.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is not built: run `make lib`")
        L = C.CDLL(path)
        L.nsp_gen_rmat_edges.restype = C.c_int
        L.nsp_gen_rmat_edges.argtypes = [C.c_int, C.c_longlong, C.c_ulonglong, C.c_void_p, C.c_void_p]
        _GEN_LIB = L
    return _GEN_LIB


def rmat_edges(scale: int, n_edges: int, seed: int):
    L = _gen_lib()
    src = np.empty(n_edges, dtype=np.int64)
    dst = np.empty(n_edges, dtype=np.int64)
    rc = L.nsp_gen_rmat_edges(scale, n_edges, seed, src.ctypes.data_as(C.c_void_p), dst.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise RuntimeError(f"nsp_gen_rmat_edges failed: {rc}")
    return src, dst


def _uniform_values(n: int, seed: int, dtype):
    """U(0,1) values from the same counter-based stream (value i depends on (seed, i) only)."""
    with np.errstate(over="ignore"):
        key = _splitmix64(np.array([seed ^ 0x5EED], dtype=np.uint64))[0]
        out = np.empty(n, dtype=dtype)
        step = 1 << 24
        for s in range(0, n, step):
            i = np.arange(s, min(n, s + step), dtype=np.uint64)
            r = _splitmix64(key ^ i) >> np.uint64(40)           # 24 bits: exact in fp32
            out[s:s + len(i)] = (r.astype(np.float64) + 0.5) / float(1 << 24)
    return out


def coo_to_csr(n_rows: int, n_cols: int, src, dst, seed: int, dtype, name="", values="uniform") -> CSR:
    """Duplicates merged (kept once), self loops kept, rows column-sorted."""
    key = np.unique(src.astype(np.int64) * np.int64(n_cols) + dst.astype(np.int64))
    rows = key // n_cols
    cols = (key - rows * n_cols).astype(np.int32)
    rpt = np.zeros(n_rows + 1, dtype=np.int64)
    rpt[1:] = np.bincount(rows, minlength=n_rows)
    rpt = np.cumsum(rpt).astype(np.int32)
    if values == "uniform":
        val = _uniform_values(len(cols), seed, dtype)
    elif values == "small_int":     # exact in fp32 under any summation order
        with np.errstate(over="ignore"):
            val = ((_splitmix64(np.arange(len(cols), dtype=np.uint64) ^ np.uint64(seed)) >> np.uint64(62))
                   .astype(dtype) + 1)
    else:
        val = np.ones(len(cols), dtype=dtype)
    return CSR(n_rows, n_cols, rpt, cols, val, name)


def rmat_csr(scale: int, edge_factor: int, seed: int = 12345, dtype=np.float32, native=True,
             values="uniform") -> CSR:
    """Config C2/C4 matrix: R-MAT scale `scale`, `edge_factor` * 2^scale generated directed edges."""
    n = 1 << scale
    ne = edge_factor * n
    src, dst = (rmat_edges if native else rmat_edges_numpy)(scale, ne, seed)
    return coo_to_csr(n, n, src, dst, seed, dtype, f"rmat_s{scale}_ef{edge_factor}", values)


def laplacian5_csr(nx: int, ny: int | None = None, dtype=np.float64) -> CSR:
    """Config C3: 5-point Laplacian on an nx x ny grid, natural ordering, Dirichlet."""
    ny = nx if ny is None else ny
    n = nx * ny
    idx = np.arange(n, dtype=np.int64)
    ix, iy = idx % nx, idx // nx
    cols = np.stack([idx - nx, idx - 1, idx, idx + 1, idx + nx], axis=1)
    ok = np.stack([iy > 0, ix > 0, np.ones(n, bool), ix < nx - 1, iy < ny - 1], axis=1)
    vals = np.broadcast_to(np.array([-1, -1, 4, -1, -1], dtype=dtype), (n, 5))
    rpt = np.zeros(n + 1, dtype=np.int64)
    rpt[1:] = np.cumsum(ok.sum(axis=1))
    return CSR(n, n, rpt.astype(np.int32), cols[ok].astype(np.int32), vals[ok].copy(), f"laplace5_{nx}x{ny}")


def er_csr(n_rows: int, n_cols: int, nnz_per_row: int, seed: int = 54321, dtype=np.float64,
           values="uniform") -> CSR:
    """Uniform-random rows with exactly `nnz_per_row` distinct sorted columns (B of configs C4/C5)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    cols = rng.integers(0, n_cols, size=(n_rows, nnz_per_row), dtype=np.int64)
    cols.sort(axis=1)
    # make duplicates within a row distinct by nudging (keeps the row sorted)
    for _ in range(nnz_per_row):
        dup = cols[:, 1:] <= cols[:, :-1]
        if not dup.any():
            break
        cols[:, 1:] = np.where(dup, cols[:, :-1] + 1, cols[:, 1:])
    cols = np.minimum(cols, n_cols - 1)
    # a clipped tail can still collide: fall back to unique per row for those rows
    bad = np.flatnonzero((cols[:, 1:] <= cols[:, :-1]).any(axis=1))
    for r in bad:
        cols[r] = np.sort(rng.choice(n_cols, size=nnz_per_row, replace=False))
    rpt = (np.arange(n_rows + 1, dtype=np.int64) * nnz_per_row).astype(np.int32)
    n = n_rows * nnz_per_row
    if values == "uniform":
        val = _uniform_values(n, seed, dtype)
    else:
        val = np.ones(n, dtype=dtype)
    return CSR(n_rows, n_cols, rpt, cols.reshape(-1).astype(np.int32), val, f"er_{n_rows}x{n_cols}_k{nnz_per_row}")


def powerlaw_csr(n: int, mean_nnz: int = 64, max_row: int = 65536, seed: int = 777, dtype=np.float64,
                 values="uniform") -> CSR:
    """Config C5: row lengths Pareto-truncated to [1, max_row] with mean ~mean_nnz and at least one
    row of exactly max_row; uniform random distinct sorted columns."""
    rng = np.random.Generator(np.random.PCG64(seed))
    alpha = 1.5
    raw = (rng.pareto(alpha, size=n) + 1.0)
    raw = np.minimum(raw, float(max_row))
    scale = mean_nnz / raw.mean()
    lens = np.clip(np.rint(raw * scale), 1, min(max_row, n)).astype(np.int64)
    lens[int(rng.integers(0, n))] = min(max_row, n)
    rpt = np.zeros(n + 1, dtype=np.int64)
    rpt[1:] = np.cumsum(lens)
    total = int(rpt[-1])
    assert total < 2 ** 31, "power-law matrix exceeds int32 nnz"
    # sorted distinct columns per row: stratified sampling (one column per equal-width stratum)
    row_of = np.repeat(np.arange(n, dtype=np.int64), lens)
    k = np.arange(total, dtype=np.int64) - rpt[row_of]
    # integer strata [k*n/len, (k+1)*n/len): non-empty because len <= n, disjoint and ordered, so
    # one draw per stratum gives strictly increasing columns
    ln = lens[row_of]
    base = (k * n) // ln
    size = ((k + 1) * n) // ln - base
    cols = base + np.minimum((rng.random(total) * size).astype(np.int64), size - 1)
    if values == "uniform":
        val = _uniform_values(total, seed, dtype)
    else:
        val = np.ones(total, dtype=dtype)
    return CSR(n, n, rpt.astype(np.int32), cols.astype(np.int32), val, f"powerlaw_{n}_m{mean_nnz}")
