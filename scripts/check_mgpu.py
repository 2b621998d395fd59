"""torchrun --nproc-per-node N scripts/check_mgpu.py [--scale S]: the gathered C of every gather path (tile
pusher = "fused", copy kernel, copy-engine pipeline, NCCL) must equal, on EVERY rank, the CPU oracle's product
(scale <= 14) and the single-GPU product of this library (any scale), bit for bit (exact integer values).
tests/test_multi_gpu_gpu.py runs it under pytest on boxes with more than one GPU."""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nsparse_b200 as ns  # noqa: E402
from nsparse_b200 import gen  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=int, default=16)
ap.add_argument("--push-sms", type=int, default=0)
args = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = ns.Context(local)
if args.push_sms:
    ctx.set_option("push_sms", args.push_sms)
for dtype in (np.float32, np.float64):
    a = gen.rmat_csr(args.scale, 16, seed=3, dtype=dtype, values="small_int")
    a.memcpy(local)
    ref = ns.spgemm_kernel_hash(a, a, ctx)
    ctx.sync()
    r_rpt, r_col, r_val = ref.to_host()
    r_fold = ref.fold(ctx)
    if args.scale <= 14:
        from oracle import oracle

        o_rpt, o_col, o_val = oracle.spgemm(a.rpt, a.col, a.val, a.rpt, a.col, a.val, acc_double=True, n_cols=a.N)
        assert np.array_equal(o_rpt, r_rpt) and np.array_equal(o_col, r_col) and np.array_equal(o_val, r_val), "single GPU vs oracle"
    cuts, total_ip = ns.partition_rows_by_ip(a.rpt, a.col, a.rpt, world)
    a_loc = ns.row_block(a, cuts[rank], cuts[rank + 1]).memcpy(local)
    peers = ns.PeerBuffers(ctx, fused=False, pieces=0)
    fused = ns.PeerBuffers(ctx, fused=True)
    piped = ns.PeerBuffers(ctx, fused=False, pieces=5)
    for mode, p in (("fused", fused), ("fused again", fused), ("push", peers), ("nccl", None), ("pipelined", piped)):
        c = ns.spgemm_kernel_hash_mgpu(a_loc, a, cuts, a.M, total_ip, ctx, peers=p)
        g_rpt, g_col, g_val = c.to_host()
        ok = c.nnz == ref.nnz and np.array_equal(g_rpt, r_rpt) and np.array_equal(g_col, r_col) and np.array_equal(g_val, r_val)
        ok = ok and c.fold(ctx)[:2] == r_fold[:2]
        print(f"rank {rank} {np.dtype(dtype).name} {mode}: nnz={c.nnz} {'OK' if ok else 'MISMATCH'}", flush=True)
        assert ok
        del c
    for p in (peers, fused, piped):
        p.release()
dist.barrier()
dist.destroy_process_group()
