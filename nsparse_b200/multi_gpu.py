"""Multi-GPU SpGEMM: 1-D row blocking of A, B replicated, allgatherv of C's row blocks.

New relative to the reference (single GPU, no communication code at all).
One process per GPU (torch.distributed, NCCL over NVLink/NVSwitch); rows of C are independent given
all of B (row-wise Gustavson), so each rank runs the unchanged single-GPU pipeline on its block and
the only exchange is the final gather.  NCCL has no native v-collective: the gather of
unequal blocks is one broadcast per rank and array, straight into the final buffers at their
displacements (no staging, no padding to the largest block).

The host logic (partition, displacement arithmetic, row-pointer rebasing) is backend agnostic and is
covered on CPU with the gloo backend (tests/test_multi_gpu_host.py).
"""
from __future__ import annotations

import numpy as np

from .csr import CSR, DeviceCSR64


def partition_rows_by_ip(a_rpt, a_col, b_rpt, nparts: int):
    """Cut points of `nparts` contiguous row blocks of A with ~equal intermediate products
    (the quantity get_spgemm_flop counts, kernel_spgemm_cu_csr.cu:18-33).  Returns (cuts, total_ip):
    block r = rows [cuts[r], cuts[r+1])."""
    a_rpt = np.asarray(a_rpt, dtype=np.int64)
    blen = np.diff(np.asarray(b_rpt, dtype=np.int64))
    per_entry = blen[np.asarray(a_col)]
    cs = np.concatenate([[0], np.cumsum(per_entry)])
    prefix = cs[a_rpt]                            # products before each row, length M + 1
    total = int(prefix[-1])
    M = len(a_rpt) - 1
    cuts = [0]
    for p in range(1, nparts):
        cuts.append(int(np.searchsorted(prefix, total * p // nparts, side="left")))
    cuts.append(M)
    for i in range(1, len(cuts)):
        cuts[i] = min(max(cuts[i], cuts[i - 1]), M)
    return cuts, total


def row_block(a: CSR, r0: int, r1: int) -> CSR:
    """Rows [r0, r1) of A as a CSR of its own (row pointer rebased to 0)."""
    lo, hi = int(a.rpt[r0]), int(a.rpt[r1])
    return CSR(r1 - r0, a.N, (a.rpt[r0:r1 + 1] - a.rpt[r0]).astype(np.int32), a.col[lo:hi], a.val[lo:hi],
               f"{a.matrix_name}[{r0}:{r1}]")


def allgatherv_csr(rpt_local, col_local, val_local, nnz_local: int, cuts, n_rows: int, group=None):
    """Gather the row blocks of C held by the ranks of `group` into the full matrix on every rank.

    rpt_local: int64 tensor [rows_local + 1] starting at 0; col_local / val_local: at least nnz_local
    entries.  Returns (rpt int64 [n_rows + 1], col, val, nnz_total).  Works on any backend / device."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = rpt_local.device
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([nnz_local], dtype=torch.int64, device=dev), group=group)
    sz = [int(s.item()) for s in sizes]
    disp = np.concatenate([[0], np.cumsum(sz)]).astype(np.int64)
    tot = int(disp[-1])
    col = torch.empty(max(tot, 1), dtype=col_local.dtype, device=dev)
    val = torch.empty(max(tot, 1), dtype=val_local.dtype, device=dev)
    rpt = torch.empty(n_rows + 1, dtype=torch.int64, device=dev)
    # every rank puts its block at its displacement of the final buffers, then block r is broadcast
    # from rank r in place (NCCL: one ncclBroadcast per block and array; nothing is staged or padded)
    lo, hi = int(disp[rank]), int(disp[rank + 1])
    col[lo:hi].copy_(col_local[:nnz_local])
    val[lo:hi].copy_(val_local[:nnz_local])
    rpt[cuts[rank]:cuts[rank + 1]].copy_(rpt_local[:-1] + lo)
    works = []
    for r in range(world):
        src = dist.get_global_rank(group, r) if group is not None else r
        for buf, a, b in ((col, disp[r], disp[r + 1]), (val, disp[r], disp[r + 1]), (rpt, cuts[r], cuts[r + 1])):
            if b > a:
                works.append(dist.broadcast(buf[int(a):int(b)], src=src, group=group, async_op=True))
    for w in works:
        w.wait()
    rpt[n_rows] = tot
    return rpt, col, val, tot


def spgemm_kernel_hash_mgpu(a_local: CSR, b: CSR, cuts, n_rows: int, total_ip: int, ctx=None, group=None) -> DeviceCSR64:
    """This rank's block through the single-GPU pipeline, then the gather.  a_local and b must have
    been memcpy()'d to this rank's GPU."""
    from .spgemm import spgemm_kernel_hash

    c = spgemm_kernel_hash(a_local, b, ctx)
    rpt, col, val, tot = allgatherv_csr(c.d_rpt64, c.d_col, c.d_val, c.nnz, cuts, n_rows, group)
    return DeviceCSR64(n_rows, b.N, rpt, col, val, tot, total_ip)
