"""CSR container mirroring sfCSR (cuda-c/inc/nsparse.h:62-75): host arrays rpt/col/val and device
mirrors d_rpt/d_col/d_val (torch tensors own the device memory; the native library only sees their
addresses)."""
from __future__ import annotations

import numpy as np


class CSR:
    def __init__(self, M, N, rpt, col, val, name=""):
        self.M, self.N = int(M), int(N)
        self.rpt = np.ascontiguousarray(rpt, dtype=np.int32)
        self.col = np.ascontiguousarray(col, dtype=np.int32)
        self.val = np.ascontiguousarray(val)
        assert self.val.dtype in (np.float32, np.float64)
        assert len(self.rpt) == self.M + 1 and len(self.col) == len(self.val) == int(self.rpt[-1])
        self.nnz = int(self.rpt[-1])
        self.nnz_max = int(np.diff(self.rpt).max()) if self.M else 0
        self.matrix_name = name
        self.d_rpt = self.d_col = self.d_val = None

    @property
    def dtype(self):
        return self.val.dtype

    # csr_memcpy (nsparse.cu:146-156)
    def memcpy(self, device=0, pinned=False):
        import torch

        dev = torch.device("cuda", device)
        f = (lambda a: torch.from_numpy(a).pin_memory()) if pinned else torch.from_numpy
        self.d_rpt = f(self.rpt).to(dev, non_blocking=pinned)
        self.d_col = f(self.col).to(dev, non_blocking=pinned)
        self.d_val = f(self.val).to(dev, non_blocking=pinned)
        return self

    # release_csr (nsparse.cu:209-214)
    def release(self):
        self.d_rpt = self.d_col = self.d_val = None

    def to_scipy(self):
        import scipy.sparse as sp

        return sp.csr_matrix((self.val, self.col, self.rpt), shape=(self.M, self.N))

    @staticmethod
    def from_scipy(m, dtype=None, name=""):
        m = m.tocsr()
        m.sort_indices()
        val = m.data if dtype is None else m.data.astype(dtype)
        return CSR(m.shape[0], m.shape[1], m.indptr, m.indices, val, name)


class DeviceCSR64:
    """Product of the SpGEMM: device CSR with an int64 row pointer (nnz may exceed 2^31)."""

    def __init__(self, M, N, d_rpt64, d_col, d_val, nnz, intprod):
        self.M, self.N, self.nnz, self.intprod = M, N, int(nnz), int(intprod)
        self.d_rpt64, self.d_col, self.d_val = d_rpt64, d_col, d_val

    def fold(self, ctx):
        """(hash of rpt, hash of col, sum of val, column-weighted sum of val) computed on the device
        (nsp_csr_fold_*): the hashes are exact, the sums equal up to fp64 rounding between runs."""
        import ctypes as C

        import torch

        h, f = (C.c_ulonglong * 2)(), (C.c_double * 2)()
        fn = ctx.lib.nsp_csr_fold_d if self.d_val.dtype == torch.float64 else ctx.lib.nsp_csr_fold_s
        ctx.use_torch_stream()
        ctx.check(fn(ctx.handle, self.M, self.nnz, C.c_void_p(self.d_rpt64.data_ptr()), C.c_void_p(self.d_col.data_ptr()),
                     C.c_void_p(self.d_val.data_ptr()), C.byref(h), C.byref(f)))
        return int(h[0]), int(h[1]), float(f[0]), float(f[1])

    # csr_memcpyDtH (nsparse.cu:158-168)
    def to_host(self):
        return (self.d_rpt64.cpu().numpy(), self.d_col[: self.nnz].cpu().numpy(),
                self.d_val[: self.nnz].cpu().numpy())
