"""Hash SpGEMM C = A*B -- host-side mirror of spgemm_kernel_hash (kernel_spgemm_hash_d.cu:1035-1075)
and get_spgemm_flop (kernel_spgemm_cu_csr.cu:35-57) over the C ABI."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .context import Context, default_context
from .csr import CSR, DeviceCSR64


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def get_spgemm_flop(a: CSR, b: CSR, ctx: Context | None = None) -> int:
    ctx = ctx or default_context(a.d_rpt.device.index)
    flop = C.c_longlong()
    ctx.check(ctx.lib.nsp_spgemm_flop(ctx.handle, a.M, _ptr(a.d_rpt), _ptr(a.d_col), _ptr(b.d_rpt), C.byref(flop)))
    return int(flop.value)


def spgemm_symbolic(a: CSR, b: CSR, ctx: Context | None = None):
    """Symbolic phase: returns (d_rpt64 tensor [M+1], nnz(C), intermediate products)."""
    import torch

    ctx = ctx or default_context(a.d_rpt.device.index)
    if a.N != b.M:
        raise ValueError(f"shape mismatch: A is {a.M}x{a.N}, B is {b.M}x{b.N}")
    ctx.use_torch_stream()
    d_rpt64 = torch.empty(a.M + 1, dtype=torch.int64, device=a.d_rpt.device)
    nnz, ip = C.c_longlong(), C.c_longlong()
    ctx.check(ctx.lib.nsp_spgemm_symbolic(ctx.handle, a.M, a.N, b.N, _ptr(a.d_rpt), _ptr(a.d_col),
                                          _ptr(b.d_rpt), _ptr(b.d_col), _ptr(d_rpt64),
                                          C.byref(nnz), C.byref(ip)))
    return d_rpt64, int(nnz.value), int(ip.value)


def spgemm_numeric(a: CSR, b: CSR, d_rpt64, nnz: int, ctx: Context | None = None, out=None, rows=None):
    """rows = (row0, nrows): only that range of rows is computed (the others are left untouched)."""
    import torch

    ctx = ctx or default_context(a.d_rpt.device.index)
    if a.dtype != b.dtype:
        raise ValueError("A and B must have the same precision")
    tdt = torch.float64 if a.dtype == np.float64 else torch.float32
    if out is None:
        d_col = torch.empty(max(nnz, 1), dtype=torch.int32, device=a.d_rpt.device)
        d_val = torch.empty(max(nnz, 1), dtype=tdt, device=a.d_rpt.device)
    else:
        d_col, d_val = out
        assert d_col.numel() >= nnz and d_val.numel() >= nnz and d_val.dtype == tdt
    if rows is not None:
        fn = ctx.lib.nsp_spgemm_numeric_rows_d if a.dtype == np.float64 else ctx.lib.nsp_spgemm_numeric_rows_s
        ctx.check(fn(ctx.handle, a.M, a.N, b.N, int(rows[0]), int(rows[1]), _ptr(a.d_rpt), _ptr(a.d_col), _ptr(a.d_val),
                     _ptr(b.d_rpt), _ptr(b.d_col), _ptr(b.d_val), _ptr(d_rpt64), _ptr(d_col), _ptr(d_val)))
        return d_col, d_val
    fn = ctx.lib.nsp_spgemm_numeric_d if a.dtype == np.float64 else ctx.lib.nsp_spgemm_numeric_s
    ctx.check(fn(ctx.handle, a.M, a.N, b.N, _ptr(a.d_rpt), _ptr(a.d_col), _ptr(a.d_val),
                 _ptr(b.d_rpt), _ptr(b.d_col), _ptr(b.d_val), _ptr(d_rpt64), _ptr(d_col), _ptr(d_val)))
    return d_col, d_val


def spgemm_kernel_hash(a: CSR, b: CSR, ctx: Context | None = None, out=None) -> DeviceCSR64:
    """C = A*B on the device.  a, b must have been memcpy()'d.  Like the reference call the result
    lives on the device; unlike it the row pointer is int64 and the call returns without a full
    device sync (the symbolic phase syncs once to learn nnz(C))."""
    d_rpt64, nnz, ip = spgemm_symbolic(a, b, ctx)
    d_col, d_val = spgemm_numeric(a, b, d_rpt64, nnz, ctx, out)
    return DeviceCSR64(a.M, b.N, d_rpt64, d_col, d_val, nnz, ip)
