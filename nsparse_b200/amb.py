"""AMB SpMV -- host-side mirror of sf_csr2amb / sf_spmv_amb / init_plan / set_plan
(cuda-c/src/conversion/convert_amb.cu:835-929, cuda-c/src/kernel/kernel_spmv_amb.cu:98-104,
cuda-c/src/nsparse.cu:171-187) over the C ABI."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .context import Context, default_context
from .csr import CSR

USHORT_MAX = 65536
MAX_BLOCK_SIZE = 20


class Plan:
    """sfPlan (nsparse.h:50-59)."""

    def __init__(self):
        # init_plan (nsparse.cu:171-174)
        self.isPlan = False
        self.seg_size = 0
        self.seg_num = 0
        self.block_size = 0
        self.thread_grid = 0
        self.thread_block = 0

    def set_plan(self, seg_size: int, block_size: int):
        """set_plan (nsparse.cu:176-187): seg_size clamped to 65536, block_size outside [1,20] -> 1."""
        self.isPlan = True
        self.seg_size = min(int(seg_size), USHORT_MAX)
        self.block_size = int(block_size) if 1 <= int(block_size) <= MAX_BLOCK_SIZE else 1
        return self


class AMB:
    """sfAMB (nsparse.h:78-107): the device arrays live in the native nsp_amb struct."""

    def __init__(self, c_amb: _lib.nsp_amb, dtype, ctx: Context):
        self._c = c_amb
        self.dtype = np.dtype(dtype)
        self.ctx = ctx

    def __getattr__(self, name):
        c = self.__dict__.get("_c")
        if c is not None and name in {f[0] for f in _lib.nsp_amb._fields_}:
            return getattr(c, name)
        raise AttributeError(name)

    def _fetch(self, ptr, count, dt):
        out = np.empty(count, dtype=dt)
        if count:
            self.ctx.check(self.ctx.lib.nsp_memcpy_d2h(self.ctx.handle, out.ctypes.data_as(C.c_void_p),
                                                       C.c_void_p(ptr), out.nbytes))
        return out

    def to_host(self) -> dict:
        """All arrays and scalars, keyed like oracle/amb.py's convert_amb result."""
        c = self._c
        lanes = c.c_size * 32
        return dict(
            M=c.M, N=c.N, pad_M=c.pad_M, chunk=c.chunk, SIGMA=c.SIGMA, seg_size=int(c.seg_size),
            seg_num=int(c.seg_num), group_num_col=int(c.seg_num), block_size=c.block_size, c_size=c.c_size,
            nnz=c.nnz,
            cs=self._fetch(c.d_cs, c.c_size, np.int32), cl=self._fetch(c.d_cl, c.c_size, np.uint32),
            sellcs_col=self._fetch(c.d_sellcs_col, c.nnz // max(c.block_size, 1), np.uint16),
            sellcs_val=self._fetch(c.d_sellcs_val, c.nnz, self.dtype),
            s_write_permutation=self._fetch(c.d_s_write_permutation, lanes, np.uint16),
            s_write_permutation_offset=self._fetch(c.d_s_write_permutation_offset, c.c_size, np.uint16),
            write_permutation=self._fetch(c.d_write_permutation, lanes, np.int32),
        )

    # release_amb (nsparse.cu:226-235)
    def release(self):
        if self._c is not None and self.ctx.handle:
            self.ctx.lib.nsp_amb_free(self.ctx.handle, C.byref(self._c))
        self._c = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


def csr2amb(csr: CSR, plan: Plan | None = None, x=None, autotune: bool = False,
            ctx: Context | None = None) -> AMB:
    """sf_csr2amb.  plan.isPlan: use its seg_size / block_size; otherwise they are chosen (footprint
    model, or by timing when autotune=True, which needs the device vector x) and written back to the plan."""
    ctx = ctx or default_context(csr.d_rpt.device.index)
    ctx.use_torch_stream()
    seg, bs = (plan.seg_size, plan.block_size) if (plan is not None and plan.isPlan) else (0, 0)
    fn = ctx.lib.nsp_csr2amb_d if csr.dtype == np.float64 else ctx.lib.nsp_csr2amb_s
    c = _lib.nsp_amb()
    px = C.c_void_p(x.data_ptr()) if x is not None else C.c_void_p(0)
    ctx.check(fn(ctx.handle, csr.M, csr.N, csr.nnz, C.c_void_p(csr.d_rpt.data_ptr()),
                 C.c_void_p(csr.d_col.data_ptr()), C.c_void_p(csr.d_val.data_ptr()),
                 int(seg), int(bs), 1 if autotune else 0, px, C.byref(c)))
    if plan is not None:
        plan.isPlan = True
        plan.seg_size, plan.seg_num, plan.block_size = int(c.seg_size), int(c.seg_num), c.block_size
        plan.thread_grid, plan.thread_block = int(c.thread_grid), int(c.thread_block)
    return AMB(c, csr.dtype, ctx)


def spmv_amb(amb: AMB, x, out=None, ctx: Context | None = None):
    """sf_spmv_amb: y = A x on the device (torch tensors)."""
    import torch

    ctx = ctx or amb.ctx
    ctx.use_torch_stream()
    tdt = torch.float64 if amb.dtype == np.float64 else torch.float32
    assert x.dtype == tdt and x.numel() >= amb.N
    y = out if out is not None else torch.empty(amb.M, dtype=tdt, device=x.device)
    fn = ctx.lib.nsp_spmv_amb_d if amb.dtype == np.float64 else ctx.lib.nsp_spmv_amb_s
    ctx.check(fn(ctx.handle, C.byref(amb._c), C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr())))
    return y
