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


def partition_rows_by_cost(a_rpt, a_col, b_rpt, c_rpt, nparts: int, nnz_weight: float):
    """Cut points balancing  intermediate products + nnz_weight * nnz(C_i)  per block.  With the product
    gathered on every GPU a rank's time is its compute (~ products) plus the bytes it has to send
    (~ its entries of C, times the number of peers), and on power-law inputs the two are distributed very
    differently over the rows: the hub rows at the top compress 16:1, the tail 2:1.  c_rpt is the row pointer
    of C from a previous symbolic phase (any product on the same pattern)."""
    a_rpt = np.asarray(a_rpt, dtype=np.int64)
    blen = np.diff(np.asarray(b_rpt, dtype=np.int64))
    cs = np.concatenate([[0], np.cumsum(blen[np.asarray(a_col)])])
    ip_prefix = cs[a_rpt].astype(np.float64)
    prefix = ip_prefix + nnz_weight * np.asarray(c_rpt, dtype=np.float64)
    M = len(a_rpt) - 1
    cuts = [0]
    for p in range(1, nparts):
        cuts.append(int(np.searchsorted(prefix, prefix[-1] * p / nparts, side="left")))
    cuts.append(M)
    for i in range(1, len(cuts)):
        cuts[i] = min(max(cuts[i], cuts[i - 1]), M)
    return cuts, int(ip_prefix[-1])


def partition_rows_by_measured(a_rpt, a_col, b_rpt, old_cuts, seconds, nparts: int):
    """Feedback cut for REPEATED products on one pattern: `seconds[r]` is what rank r needed for its block
    [old_cuts[r], old_cuts[r+1]) of the previous product.  Every row is charged its intermediate products divided by
    the speed (products per second) its old block was computed at, and the new cuts split that cost evenly.
    Power-law inputs need it: the hub rows at the top are computed at a different rate per product than the tail
    (more windows and chunks per row, the rows with more than 1024 entries of A take the red.global path)."""
    a_rpt = np.asarray(a_rpt, dtype=np.int64)
    blen = np.diff(np.asarray(b_rpt, dtype=np.int64))
    cs = np.concatenate([[0], np.cumsum(blen[np.asarray(a_col)])])
    ip_prefix = cs[a_rpt].astype(np.float64)                  # products before each row
    M = len(a_rpt) - 1
    cost = np.zeros(M + 1, dtype=np.float64)
    for r in range(len(old_cuts) - 1):
        lo, hi = old_cuts[r], old_cuts[r + 1]
        ip_r = ip_prefix[hi] - ip_prefix[lo]
        per_product = (seconds[r] / ip_r) if ip_r > 0 else 0.0
        cost[lo + 1:hi + 1] = cost[lo] + (ip_prefix[lo + 1:hi + 1] - ip_prefix[lo]) * per_product
    cuts = [0]
    for p in range(1, nparts):
        cuts.append(int(np.searchsorted(cost, cost[-1] * p / nparts, side="left")))
    cuts.append(M)
    for i in range(1, len(cuts)):
        cuts[i] = min(max(cuts[i], cuts[i - 1]), M)
    return cuts, int(ip_prefix[-1])


def partition_rows_minmax(a_rpt, a_col, b_rpt, c_rpt, old_cuts, seconds, nparts: int, bytes_per_entry: int,
                          out_gbs: float = 400.0, tail_gbs: float | None = None):
    """Feedback cut that bounds BOTH what a rank computes and what it sends.  With the product gathered on every GPU a
    rank's block costs its compute time -- as in partition_rows_by_measured: products at the rate the row's old block
    was computed at -- plus whatever of its outbound transfer (its entries of C * bytes_per_entry * (nparts - 1)) is
    left when the kernels end: `out_gbs` leave while they run (what the copy engines of one B200 sustain towards 7
    peers, profiles/r2_bench_c2_gpus8_*.json), the rest at `tail_gbs` (copy engines and SM stores together,
    csrc/peer_dma.cu; None: the copy engines alone, i.e. max(compute, transfer)).  The cuts minimise the largest cost
    over all blocks (binary search on the bound, greedy blocks).  At 2 GPUs the compute is the active bound, at 8 the
    transfer: balancing compute alone left one GPU sending 118 GB of the 78 GB product (2.1e9 of its entries to 7
    peers) while another sent 22 GB."""
    a_rpt = np.asarray(a_rpt, dtype=np.int64)
    blen = np.diff(np.asarray(b_rpt, dtype=np.int64))
    cs = np.concatenate([[0], np.cumsum(blen[np.asarray(a_col)])])
    ip_prefix = cs[a_rpt].astype(np.float64)
    M = len(a_rpt) - 1
    comp = np.zeros(M + 1, dtype=np.float64)
    for r in range(len(old_cuts) - 1):
        lo, hi = old_cuts[r], old_cuts[r + 1]
        ip_r = ip_prefix[hi] - ip_prefix[lo]
        per_product = (seconds[r] / ip_r) if ip_r > 0 else 0.0
        comp[lo + 1:hi + 1] = comp[lo] + (ip_prefix[lo + 1:hi + 1] - ip_prefix[lo]) * per_product
    out_bytes = np.asarray(c_rpt, dtype=np.float64) * (bytes_per_entry * max(nparts - 1, 0))
    during, after = out_gbs * 1e9, (tail_gbs if tail_gbs else out_gbs) * 1e9

    def cost(start, end):
        t = comp[end] - comp[start]
        return t + max(0.0, out_bytes[end] - out_bytes[start] - during * t) / after

    def blocks_for(T):
        cuts, start = [0], 0
        while start < M and len(cuts) <= nparts:
            lo, hi = start + 1, M                      # a single row always fits (it cannot be split)
            while lo < hi:
                mid = (lo + hi + 1) // 2
                if cost(start, mid) <= T:
                    lo = mid
                else:
                    hi = mid - 1
            cuts.append(lo)
            start = lo
        return cuts

    lo_t, hi_t = 0.0, cost(0, M) + 1e-9
    for _ in range(50):
        mid = 0.5 * (lo_t + hi_t)
        c = blocks_for(mid)
        if c[-1] >= M and len(c) - 1 <= nparts:
            hi_t = mid
        else:
            lo_t = mid
    cuts = blocks_for(hi_t)
    cuts = cuts[:nparts] + [M] if len(cuts) > nparts else cuts + [M] * (nparts + 1 - len(cuts))
    for i in range(1, len(cuts)):
        cuts[i] = min(max(cuts[i], cuts[i - 1]), M)
    return cuts, int(ip_prefix[-1])


def partition_rows_by_ip_device(a, b, nparts: int):
    """partition_rows_by_ip for matrices that live on the GPU (nsparse_b200.gen.DeviceCSR): the same cuts, computed
    with torch on the device.  Returns (cuts, total_ip)."""
    import torch

    blen = (b.d_rpt[1:] - b.d_rpt[:-1]).long()
    cs = torch.cumsum(blen[a.d_col.long()], 0)
    prefix = torch.zeros(a.M + 1, dtype=torch.int64, device=cs.device)
    rp = a.d_rpt.long()
    nz = rp[1:] > 0
    prefix[1:][nz] = cs[rp[1:][nz] - 1]
    del cs
    total = int(prefix[-1])
    targets = torch.tensor([total * p // nparts for p in range(1, nparts)], dtype=torch.int64, device=prefix.device)
    mid = torch.searchsorted(prefix, targets, right=False).tolist() if nparts > 1 else []
    cuts = [0] + [int(x) for x in mid] + [a.M]
    for i in range(1, len(cuts)):
        cuts[i] = min(max(cuts[i], cuts[i - 1]), a.M)
    return cuts, total


def row_block(a, r0: int, r1: int):
    """Rows [r0, r1) of A as a CSR of its own (row pointer rebased to 0); host CSR or DeviceCSR."""
    if hasattr(a, "row_block"):
        return a.row_block(r0, r1)
    return _row_block_host(a, r0, r1)


def _row_block_host(a: CSR, r0: int, r1: int) -> CSR:
    lo, hi = int(a.rpt[r0]), int(a.rpt[r1])
    return CSR(r1 - r0, a.N, (a.rpt[r0:r1 + 1] - a.rpt[r0]).astype(np.int32), a.col[lo:hi], a.val[lo:hi],
               f"{a.matrix_name}[{r0}:{r1}]")


def _gather_sizes(nnz_local: int, dev, group=None):
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([nnz_local], dtype=torch.int64, device=dev), group=group)
    sz = [int(s.item()) for s in sizes]
    return np.concatenate([[0], np.cumsum(sz)]).astype(np.int64)


def _broadcast_blocks(rpt, col, val, disp, cuts, group=None):
    """Block r of every array is broadcast from rank r in place (NCCL: one ncclBroadcast per block and
    array; nothing is staged or padded)."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    works = []
    for r in range(world):
        src = dist.get_global_rank(group, r) if group is not None else r
        for buf, a, b in ((col, disp[r], disp[r + 1]), (val, disp[r], disp[r + 1]), (rpt, cuts[r], cuts[r + 1])):
            if b > a:
                works.append(dist.broadcast(buf[int(a):int(b)], src=src, group=group, async_op=True))
    for w in works:
        w.wait()


def allgatherv_csr(rpt_local, col_local, val_local, nnz_local: int, cuts, n_rows: int, group=None):
    """Gather the row blocks of C held by the ranks of `group` into the full matrix on every rank.

    rpt_local: int64 tensor [rows_local + 1] starting at 0; col_local / val_local: at least nnz_local
    entries.  Returns (rpt int64 [n_rows + 1], col, val, nnz_total).  Works on any backend / device."""
    import torch
    import torch.distributed as dist

    rank = dist.get_rank(group)
    dev = rpt_local.device
    disp = _gather_sizes(nnz_local, dev, group)
    tot = int(disp[-1])
    col = torch.empty(max(tot, 1), dtype=col_local.dtype, device=dev)
    val = torch.empty(max(tot, 1), dtype=val_local.dtype, device=dev)
    rpt = torch.empty(n_rows + 1, dtype=torch.int64, device=dev)
    lo, hi = int(disp[rank]), int(disp[rank + 1])
    col[lo:hi].copy_(col_local[:nnz_local])
    val[lo:hi].copy_(val_local[:nnz_local])
    rpt[cuts[rank]:cuts[rank + 1]].copy_(rpt_local[:-1] + lo)
    _broadcast_blocks(rpt, col, val, disp, cuts, group)
    rpt[n_rows] = tot
    return rpt, col, val, tot


class _DeviceArray:
    """Raw device memory as something torch.as_tensor understands (the memory stays owned by PeerBuffers)."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3,
                                         "strides": None}


class PeerBuffers:
    """The full-size C arrays of every rank, mapped into this process with CUDA IPC (nsp_peer_alloc /
    nsp_peer_open), so that a rank can store its row block straight into every other GPU's copy over
    NVLink / NVSwitch.  Grow-only: handles are exchanged again only when a product needs more room."""

    _TYPES = {"col": ("<i4", 4), "rpt": ("<i8", 8)}

    def __init__(self, ctx, group=None, fused=True, pieces=4):
        """How a rank's block reaches the other GPUs:
          fused (default)        the numeric kernels store every chunk of C into the peers themselves
                                 (nsp_spgemm_set_peers);
          pieces >= 1            the numeric phase runs piece by piece (contiguous row ranges of ~equal
                                 output) and every finished piece is copied to each peer by the copy engines
                                 on side streams while the next piece is computed (nsp_copy_async);
          pieces == 0            one copy kernel after the numeric phase (nsp_push_to_peers).
        Measured on 2 x B200, R-MAT scale 20 A^2 (39 GB per direction): fused 150 ms per product, 4 pieces
        156 ms (contiguous row pieces lose the heaviest-first balance of the row queue), copy kernel 177 ms,
        one copy-engine pass 204 ms, NCCL broadcasts 252 ms; one GPU: 206 ms."""
        self.ctx, self.group, self.fused, self.pieces = ctx, group, fused, pieces
        self._streams = None
        self.cap_nnz, self.n_rows, self.dtype = -1, -1, None
        self.col = self.val = self.rpt = None
        self._own, self._opened = {}, []
        self.last_compute_s = 0.0

    def _others(self, key):
        import torch.distributed as dist

        rank = dist.get_rank(self.group)
        return [p for r, p in enumerate(self.ptrs[key]) if r != rank]

    def copy_streams(self, dev):
        import torch

        n = len(self._others("col"))
        if self._streams is None or len(self._streams) != n:
            self._streams = [torch.cuda.Stream(device=dev) for _ in range(n)]
        return self._streams

    def copy_to_peers(self, key: str, src, elem_offset: int, event, dev):
        """Copy-engine transfer of `src` (a slice of this rank's array `key` starting at elem_offset) to the
        same place on every other rank, each on its own stream, after `event`."""
        import ctypes as C

        if src.numel() == 0:
            return
        es = src.element_size()
        for st, base in zip(self.copy_streams(dev), self._others(key)):
            st.wait_event(event)
            self.ctx.check(self.ctx.lib.nsp_copy_async(self.ctx.handle, C.c_void_p(base + elem_offset * es),
                                                       C.c_void_p(src.data_ptr()), src.numel() * es,
                                                       C.c_void_p(st.cuda_stream)))

    def set_fused_targets(self, elem_offset: int):
        import ctypes as C

        cols, vals = self._others("col"), self._others("val")
        if len(cols) > 7:
            raise ValueError("fused peer stores support at most 8 GPUs")
        self.ctx.check(self.ctx.lib.nsp_spgemm_set_peers(self.ctx.handle, len(cols), (C.c_void_p * len(cols))(*cols),
                                                         (C.c_void_p * len(vals))(*vals), elem_offset))

    def check_status(self):
        import ctypes as C

        err = C.c_int(0)
        self.ctx.check(self.ctx.lib.nsp_spgemm_peers_status(self.ctx.handle, C.byref(err)))

    def clear_fused_targets(self):
        self.ctx.check(self.ctx.lib.nsp_spgemm_set_peers(self.ctx.handle, 0, None, None, 0))

    def release(self):
        """Collective.  Two phases with a barrier in between: every rank first closes the mappings it
        imported, and only when all ranks have done so are the exported buffers freed (cudaFree of an exported
        region that an importer still has open is undefined behaviour)."""
        import ctypes as C

        import torch.distributed as dist

        self.col = self.val = self.rpt = None
        for p in self._opened:
            self.ctx.lib.nsp_peer_close(self.ctx.handle, C.c_void_p(p))
        self._opened = []
        if self._own and dist.is_initialized():
            dist.barrier(self.group)
        for p in self._own.values():
            self.ctx.lib.nsp_peer_free(self.ctx.handle, C.c_void_p(p))
        self._own = {}
        self.cap_nnz = -1

    def ensure(self, tot: int, n_rows: int, tdt, dev):
        """Collective (every rank calls it with the same arguments)."""
        import ctypes as C

        import torch
        import torch.distributed as dist

        if tot <= self.cap_nnz and n_rows == self.n_rows and tdt == self.dtype:
            return
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        torch.cuda.synchronize(dev)
        dist.barrier(self.group)               # nobody still pushes into the buffers that are about to go
        self.release()
        torch.cuda.empty_cache()
        cap = max(int(tot), 4)
        vt = ("<f8", 8) if tdt == torch.float64 else ("<f4", 4)
        spec = {"col": (cap,) + self._TYPES["col"], "val": (cap,) + vt, "rpt": (n_rows + 1,) + self._TYPES["rpt"]}
        handles = {}
        for k, (n, typestr, es) in spec.items():
            p, h = C.c_void_p(), C.create_string_buffer(64)
            self.ctx.check(self.ctx.lib.nsp_peer_alloc(self.ctx.handle, n * es, C.byref(p), h))
            self._own[k] = p.value
            handles[k] = h.raw
            setattr(self, k, torch.as_tensor(_DeviceArray(p.value, n, typestr), device=dev))
        everyone = [None] * world
        dist.all_gather_object(everyone, handles, group=self.group)
        self.ptrs = {"col": [], "val": [], "rpt": []}
        for r in range(world):
            for k in ("col", "val", "rpt"):
                if r == rank:
                    self.ptrs[k].append(self._own[k])
                else:
                    p = C.c_void_p()
                    self.ctx.check(self.ctx.lib.nsp_peer_open(self.ctx.handle, everyone[r][k], C.byref(p)))
                    self._opened.append(p.value)
                    self.ptrs[k].append(p.value)
        self.cap_nnz, self.n_rows, self.dtype = cap, n_rows, tdt
        dist.barrier(self.group)

    def push(self, key: str, src, elem_offset: int):
        """Store `src` (a slice of this rank's own array `key`, starting at element elem_offset) into the
        same place of every OTHER rank's array."""
        import ctypes as C

        import torch.distributed as dist

        rank = dist.get_rank(self.group)
        dst = [p for r, p in enumerate(self.ptrs[key]) if r != rank]
        if not dst or src.numel() == 0:
            return
        arr = (C.c_void_p * len(dst))(*dst)
        es = src.element_size()
        self.ctx.check(self.ctx.lib.nsp_push_to_peers(self.ctx.handle, len(dst), arr, elem_offset * es,
                                                      C.c_void_p(src.data_ptr()), src.numel() * es))


def spgemm_kernel_hash_mgpu(a_local: CSR, b: CSR, cuts, n_rows: int, total_ip: int, ctx=None, group=None,
                            peers: "PeerBuffers | None" = None) -> DeviceCSR64:
    """This rank's block through the single-GPU pipeline, then the gather.  a_local and b must have
    been memcpy()'d to this rank's GPU.  The numeric phase writes the block straight into its place in
    the full C (the displacements are known after the symbolic phase), so a rank never holds more than
    one copy of C: 78 GB at R-MAT scale 20."""
    import torch
    import torch.distributed as dist

    from .spgemm import spgemm_numeric, spgemm_symbolic

    import time

    rank = dist.get_rank(group)
    t_start = time.perf_counter()
    d_rpt64, nnz, _ = spgemm_symbolic(a_local, b, ctx)       # (synchronises: nnz comes back to the host)
    t_own = time.perf_counter() - t_start
    dev = d_rpt64.device
    disp = _gather_sizes(nnz, dev, group)
    tot = int(disp[-1])
    tdt = torch.float64 if a_local.dtype == np.float64 else torch.float32
    if peers is not None:
        # NVLink push: the block is computed in place and stored into every peer's arrays by one kernel
        # per array (nsp_push_to_peers); the exchange of the sizes above is also the point after which no
        # rank still reads the previous product.
        peers.ensure(tot, n_rows, tdt, dev)
        lo, hi = int(disp[rank]), int(disp[rank + 1])
        col, val, rpt = peers.col, peers.val, peers.rpt
        out = (col[lo:max(hi, lo + 1)], val[lo:max(hi, lo + 1)])
        if peers.fused:
            # the numeric kernels count finished tiles of C, the pusher kernel sends them to all peers meanwhile
            peers.set_fused_targets(lo)
            t_num = time.perf_counter()
            try:
                spgemm_numeric(a_local, b, d_rpt64, nnz, ctx, out=out)
            finally:
                peers.clear_fused_targets()
            peers.check_status()
            # the rank's own compute time, for the feedback partitions: symbolic phase + numeric call up to the end of
            # its kernels (the call itself returns when the last tile has been handed to the copy engines / SMs, which
            # with many peers is most of the transfer)
            import ctypes as C

            kms = C.c_double(0.0)
            ctx.check(ctx.lib.nsp_spgemm_peers_stats(ctx.handle, None, None, C.byref(kms)))
            peers.last_compute_s = t_own + (kms.value * 1e-3 if kms.value > 0 else time.perf_counter() - t_num)
        elif peers.pieces >= 1 and a_local.M > 0:
            # pipeline: piece k+1 is computed while the copy engines carry piece k to the peers
            main = torch.cuda.current_stream(dev)
            k = min(peers.pieces, a_local.M)
            targets = torch.tensor([nnz * i // k for i in range(1, k)], dtype=torch.int64, device=dev)
            rows = [0] + torch.searchsorted(d_rpt64, targets, right=True).clamp_(0, a_local.M).tolist() + [a_local.M]
            offs = d_rpt64[torch.tensor(rows, device=dev)].tolist()
            for i in range(k):
                r0, r1 = rows[i], max(rows[i + 1], rows[i])
                if r1 > r0:
                    spgemm_numeric(a_local, b, d_rpt64, nnz, ctx, out=out, rows=(r0, r1 - r0))
                ev = torch.cuda.Event()
                ev.record(main)
                s0, s1 = lo + offs[i], lo + max(offs[i + 1], offs[i])
                peers.copy_to_peers("col", col[s0:s1], s0, ev, dev)
                peers.copy_to_peers("val", val[s0:s1], s0, ev, dev)
            for st in peers.copy_streams(dev):
                main.wait_stream(st)
        else:
            spgemm_numeric(a_local, b, d_rpt64, nnz, ctx, out=out)
            peers.push("col", col[lo:hi], lo)
            peers.push("val", val[lo:hi], lo)
        rpt[cuts[rank]:cuts[rank + 1]].copy_(d_rpt64[:-1] + lo)
        rpt[n_rows] = tot
        peers.push("rpt", rpt[cuts[rank]:cuts[rank + 1]], cuts[rank])
        torch.cuda.current_stream(dev).synchronize()
        dist.barrier(group)                      # every block of every rank has landed
        return DeviceCSR64(n_rows, b.N, rpt, col[:max(tot, 1)], val[:max(tot, 1)], tot, total_ip)
    col = torch.empty(max(tot, 1), dtype=torch.int32, device=dev)
    val = torch.empty(max(tot, 1), dtype=tdt, device=dev)
    rpt = torch.empty(n_rows + 1, dtype=torch.int64, device=dev)
    lo, hi = int(disp[rank]), int(disp[rank + 1])
    spgemm_numeric(a_local, b, d_rpt64, nnz, ctx, out=(col[lo:max(hi, lo + 1)], val[lo:max(hi, lo + 1)]))
    rpt[cuts[rank]:cuts[rank + 1]].copy_(d_rpt64[:-1] + lo)
    _broadcast_blocks(rpt, col, val, disp, cuts, group)
    rpt[n_rows] = tot
    return DeviceCSR64(n_rows, b.N, rpt, col, val, tot, total_ip)
