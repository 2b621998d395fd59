"""Development: timeline of the copy-engine gather on ONE GPU (the peer is a second buffer on the same device):
NSP_DMA_TRACE=1 python scripts/trace_dma_order.py [scale] -- with and without the per-launch profiling events."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nsparse_b200 as ns  # noqa: E402
from nsparse_b200 import gen  # noqa: E402

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 18
a = gen.rmat_csr(scale, 16, dtype=np.float32)
a.memcpy()
ctx = ns.Context(0)
d_rpt64, nnz, _ = ns.spgemm_symbolic(a, a, ctx)
col = torch.empty(nnz, dtype=torch.int32, device="cuda")
val = torch.empty(nnz, dtype=torch.float32, device="cuda")
pcol = torch.empty(nnz, dtype=torch.int32, device="cuda")
pval = torch.empty(nnz, dtype=torch.float32, device="cuda")
for prof in (0, 0, 1, 1, 0):
    ctx.profile(bool(prof))
    d_rpt64, nnz, _ = ns.spgemm_symbolic(a, a, ctx)
    ctx.check(ctx.lib.nsp_spgemm_set_peers(ctx.handle, 1, (C.c_void_p * 1)(pcol.data_ptr()), (C.c_void_p * 1)(pval.data_ptr()), 0))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ns.spgemm_numeric(a, a, d_rpt64, nnz, ctx, out=(col, val))
    e1.record()
    ctx.check(ctx.lib.nsp_spgemm_set_peers(ctx.handle, 0, None, None, 0))
    torch.cuda.synchronize()
    print(f"profile={prof}: numeric + gather {e0.elapsed_time(e1):.1f} ms", flush=True)
    if prof:
        for n, ms, *_ in ctx.profile_dump():
            print(f"    {n:20s} {ms:8.2f} ms")
assert torch.equal(col, pcol) and torch.equal(val, pval)
print("peer copy equals C")
