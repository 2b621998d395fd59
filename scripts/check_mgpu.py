"""torchrun --nproc-per-node N scripts/check_mgpu.py: the gathered C of the peer-push path and of the NCCL
path must both equal the single-GPU product on every rank (R-MAT scale 16, exact integer values)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nsparse_b200 as ns  # noqa: E402
from nsparse_b200 import gen  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = ns.Context(local)
for dtype in (np.float32, np.float64):
    a = gen.rmat_csr(16, 16, seed=3, dtype=dtype, values="small_int")
    a.memcpy(local)
    ref = ns.spgemm_kernel_hash(a, a, ctx)
    ctx.sync()
    r_rpt, r_col, r_val = ref.to_host()
    cuts, total_ip = ns.partition_rows_by_ip(a.rpt, a.col, a.rpt, world)
    a_loc = ns.row_block(a, cuts[rank], cuts[rank + 1]).memcpy(local)
    peers = ns.PeerBuffers(ctx, pieces=0)
    fused = ns.PeerBuffers(ctx, fused=True)
    piped = ns.PeerBuffers(ctx, pieces=5)
    for mode, p in (("push", peers), ("nccl", None), ("fused", fused), ("pipelined", piped), ("pipelined again", piped)):
        c = ns.spgemm_kernel_hash_mgpu(a_loc, a, cuts, a.M, total_ip, ctx, peers=p)
        g_rpt, g_col, g_val = c.to_host()
        ok = c.nnz == ref.nnz and np.array_equal(g_rpt, r_rpt) and np.array_equal(g_col, r_col) and np.array_equal(g_val, r_val)
        print(f"rank {rank} {np.dtype(dtype).name} {mode}: nnz={c.nnz} {'OK' if ok else 'MISMATCH'}", flush=True)
        assert ok
dist.barrier()
dist.destroy_process_group()
