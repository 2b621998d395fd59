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


# ---------------------------------------------------------------------------------------------
# The same inputs generated ON THE GPU with torch (configs C4 / C5 at full size: 2.7e8 and 1.1e9 entries are
# minutes of numpy per rank on the host, seconds here).  R-MAT edges and values are the same counter-based
# streams as above, bit for bit (tests/test_gen_device_gpu.py); the Erdos-Renyi and power-law matrices use the
# same constructions with counter-based draws instead of numpy's PCG64 stream, so they are a deterministic
# function of their seed on every rank but not the numpy matrices of er_csr / powerlaw_csr.
# ---------------------------------------------------------------------------------------------
class DeviceCSR:
    """CSR that lives on the GPU only (d_rpt int32, d_col int32, d_val); what the SpGEMM entry points read."""

    def __init__(self, M, N, d_rpt, d_col, d_val, name=""):
        self.M, self.N = int(M), int(N)
        self.d_rpt, self.d_col, self.d_val = d_rpt, d_col, d_val
        self.nnz = int(d_col.numel())
        self.matrix_name = name

    @property
    def dtype(self):
        import torch

        return np.dtype(np.float64 if self.d_val.dtype == torch.float64 else np.float32)

    def memcpy(self, device=0, pinned=False):
        return self

    def rows_to_host(self, rows) -> CSR:
        """The given rows as a host CSR (for oracle checks on a sample)."""
        import torch

        rows = torch.as_tensor(np.asarray(rows, dtype=np.int64), device=self.d_rpt.device)
        beg = self.d_rpt[rows].long()
        lens = self.d_rpt[rows + 1].long() - beg
        rpt = torch.zeros(len(rows) + 1, dtype=torch.int64, device=rows.device)
        rpt[1:] = torch.cumsum(lens, 0)
        idx = torch.repeat_interleave(beg - rpt[:-1], lens) + torch.arange(int(rpt[-1]), device=rows.device)
        return CSR(len(rows), self.N, rpt.cpu().numpy().astype(np.int32), self.d_col[idx].cpu().numpy(),
                   self.d_val[idx].cpu().numpy(), self.matrix_name + "[sample]")

    def to_host(self) -> CSR:
        return CSR(self.M, self.N, self.d_rpt.cpu().numpy(), self.d_col.cpu().numpy(), self.d_val.cpu().numpy(), self.matrix_name)

    def row_block(self, r0: int, r1: int) -> "DeviceCSR":
        lo, hi = int(self.d_rpt[r0]), int(self.d_rpt[r1])
        return DeviceCSR(r1 - r0, self.N, (self.d_rpt[r0:r1 + 1] - lo).contiguous(), self.d_col[lo:hi].clone(),
                         self.d_val[lo:hi].clone(), f"{self.matrix_name}[{r0}:{r1}]")


def _i64(c: int) -> int:
    return c - (1 << 64) if c >= (1 << 63) else c


def _lsr(x, k: int):
    return (x >> k) & ((1 << (64 - k)) - 1)


def _splitmix64_t(x):
    x = x + _i64(0x9E3779B97F4A7C15)
    x = (x ^ _lsr(x, 30)) * _i64(0xBF58476D1CE4E5B9)
    x = (x ^ _lsr(x, 27)) * _i64(0x94D049BB133111EB)
    return x ^ _lsr(x, 31)


def _key_t(seed: int, dev):
    import torch

    return _splitmix64_t(torch.tensor([_i64(seed & ((1 << 64) - 1))], dtype=torch.int64, device=dev))[0]


def _uniform_values_t(n: int, seed: int, tdt, dev, first: int = 0):
    import torch

    key = _key_t(seed ^ 0x5EED, dev)
    out = torch.empty(n, dtype=tdt, device=dev)
    step = 1 << 27
    for s in range(0, n, step):
        i = torch.arange(first + s, first + min(n, s + step), dtype=torch.int64, device=dev)
        r = _lsr(_splitmix64_t(key ^ i), 40)
        out[s:s + i.numel()] = ((r.double() + 0.5) / float(1 << 24)).to(tdt)
    return out


def _u01_t(key, idx):
    """53-bit uniform in [0, 1) from the counter stream."""
    return _lsr(_splitmix64_t(key ^ idx), 11).double() / float(1 << 53)


def rmat_edges_device(scale: int, n_edges: int, seed: int, dev):
    import torch

    key = _key_t(seed, dev)
    ta, tab, tabc = 2448131358, 3264175144, 4080218931
    src = torch.zeros(n_edges, dtype=torch.int64, device=dev)
    dst = torch.zeros(n_edges, dtype=torch.int64, device=dev)
    e64 = torch.arange(n_edges, dtype=torch.int64, device=dev) * 64
    for l in range(scale):
        r = _lsr(_splitmix64_t(key ^ (e64 + l)), 32)
        sbit = (r >= tab).long()
        dbit = (((r >= ta) & (r < tab)) | (r >= tabc)).long()
        src = (src << 1) | sbit
        dst = (dst << 1) | dbit
    return src, dst


def rmat_csr_device(scale: int, edge_factor: int, seed: int = 12345, dtype=np.float32, device=0) -> DeviceCSR:
    import torch

    dev = device if isinstance(device, torch.device) else torch.device("cuda", device)
    n = 1 << scale
    src, dst = rmat_edges_device(scale, edge_factor * n, seed, dev)
    key = torch.unique(src * n + dst)
    del src, dst
    rows = key // n
    cols = (key - rows * n).to(torch.int32)
    del key
    rpt = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    rpt[1:] = torch.cumsum(torch.bincount(rows, minlength=n), 0)
    del rows
    tdt = torch.float64 if np.dtype(dtype) == np.float64 else torch.float32
    val = _uniform_values_t(cols.numel(), seed, tdt, dev)
    return DeviceCSR(n, n, rpt.to(torch.int32), cols, val, f"rmat_s{scale}_ef{edge_factor}")


def er_csr_device(n_rows: int, n_cols: int, nnz_per_row: int, seed: int = 54321, dtype=np.float64, device=0) -> DeviceCSR:
    """Uniform-random rows with exactly nnz_per_row distinct sorted columns (B of configs C4 / C5)."""
    import torch

    dev = device if isinstance(device, torch.device) else torch.device("cuda", device)
    key = _key_t(seed, dev)
    idx = torch.arange(n_rows * nnz_per_row, dtype=torch.int64, device=dev)
    cols = (_lsr(_splitmix64_t(key ^ idx), 11) % n_cols).view(n_rows, nnz_per_row)
    cols, _ = torch.sort(cols, dim=1)
    for _ in range(nnz_per_row):
        dup = cols[:, 1:] <= cols[:, :-1]
        if not bool(dup.any()):
            break
        cols[:, 1:] = torch.where(dup, cols[:, :-1] + 1, cols[:, 1:])
    bad = ((cols[:, 1:] <= cols[:, :-1]).any(dim=1)) | (cols[:, -1] >= n_cols)
    if bool(bad.any()):      # the few rows that ran into the right edge: one column per stratum instead
        k = torch.arange(nnz_per_row, dtype=torch.int64, device=dev)
        cols[bad] = (k * n_cols) // nnz_per_row + cols[bad] % max(n_cols // nnz_per_row, 1)
    rpt = (torch.arange(n_rows + 1, dtype=torch.int64, device=dev) * nnz_per_row).to(torch.int32)
    tdt = torch.float64 if np.dtype(dtype) == np.float64 else torch.float32
    val = _uniform_values_t(n_rows * nnz_per_row, seed, tdt, dev)
    return DeviceCSR(n_rows, n_cols, rpt, cols.reshape(-1).to(torch.int32), val, f"er_{n_rows}x{n_cols}_k{nnz_per_row}")


def powerlaw_csr_device(n: int, mean_nnz: int = 64, max_row: int = 65536, seed: int = 777, dtype=np.float64,
                        device=0) -> DeviceCSR:
    """Config C5: row lengths Pareto(1.5)-truncated to [1, max_row] with mean ~mean_nnz and at least one row of
    exactly max_row; one uniformly drawn column per equal-width stratum of the row (distinct, sorted)."""
    import torch

    dev = device if isinstance(device, torch.device) else torch.device("cuda", device)
    key = _key_t(seed, dev)
    u = _u01_t(key, torch.arange(n, dtype=torch.int64, device=dev))
    raw = torch.clamp((1.0 - u) ** (-1.0 / 1.5), max=float(max_row))
    scale = mean_nnz / float(raw.mean())
    lens = torch.clamp(torch.round(raw * scale), 1, min(max_row, n)).long()
    lens[int(_lsr(_splitmix64_t(key ^ (1 << 40)), 11) % n)] = min(max_row, n)
    del u, raw
    rpt = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    rpt[1:] = torch.cumsum(lens, 0)
    total = int(rpt[-1])
    assert total < 2 ** 31, "power-law matrix exceeds int32 nnz"
    cols = torch.empty(total, dtype=torch.int32, device=dev)
    key2 = _key_t(seed + 1, dev)
    # in row blocks, so that the int64 temporaries stay at a few GB
    step = 1 << 20
    for r0 in range(0, n, step):
        r1 = min(n, r0 + step)
        ln = lens[r0:r1]
        lo, hi = int(rpt[r0]), int(rpt[r1])
        row_of = torch.repeat_interleave(torch.arange(r1 - r0, dtype=torch.int64, device=dev), ln)
        k = torch.arange(lo, hi, dtype=torch.int64, device=dev) - rpt[r0:r1][row_of]
        lnr = ln[row_of]
        base = (k * n) // lnr
        size = ((k + 1) * n) // lnr - base
        draw = _lsr(_splitmix64_t(key2 ^ torch.arange(lo, hi, dtype=torch.int64, device=dev)), 11) % size
        cols[lo:hi] = (base + draw).to(torch.int32)
        del row_of, k, lnr, base, size, draw
    tdt = torch.float64 if np.dtype(dtype) == np.float64 else torch.float32
    val = _uniform_values_t(total, seed, tdt, dev)
    return DeviceCSR(n, n, rpt.to(torch.int32), cols, val, f"powerlaw_{n}_m{mean_nnz}")
