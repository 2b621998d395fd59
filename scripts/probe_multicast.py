"""torchrun --nproc-per-node N scripts/probe_multicast.py: is NVLS multicast available through torch's
symmetric memory, and what does a multimem.st push reach against unicast peer stores?"""
import ctypes as C
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nsparse_b200 as ns  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import torch.distributed._symmetric_memory as symm_mem  # noqa: E402

ctx = ns.Context(local)
n = 1 << 28                      # 1 GiB of int32 per rank buffer
per = n // world
t = symm_mem.empty(n, dtype=torch.int32, device=dev)
hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
ptrs = [int(p) for p in hdl.buffer_ptrs]
if rank == 0:
    print(f"world {world}: multicast_ptr = {mc:#x}, buffer_ptrs ok = {len(ptrs) == world}", flush=True)
src = torch.full((per,), rank + 1, dtype=torch.int32, device=dev)
t.zero_()
torch.cuda.synchronize()
dist.barrier()


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    return e0.elapsed_time(e1) / reps


others = [p for r, p in enumerate(ptrs) if r != rank]
arr = (C.c_void_p * len(others))(*others)
uni = timed(lambda: ctx.check(ctx.lib.nsp_push_to_peers(ctx.handle, len(others), arr, rank * per * 4,
                                                        C.c_void_p(src.data_ptr()), per * 4)))
msg = f"rank {rank}: unicast stores to {len(others)} peers: {uni:.2f} ms, {per * 4 * len(others) / uni / 1e6:.0f} GB/s out"
if mc:
    t.zero_()
    torch.cuda.synchronize()
    dist.barrier()
    mul = timed(lambda: ctx.check(ctx.lib.nsp_push_multicast(ctx.handle, C.c_void_p(mc), rank * per * 4,
                                                             C.c_void_p(src.data_ptr()), per * 4)))
    torch.cuda.synchronize()
    dist.barrier()
    ok = all(bool((t[r * per:(r + 1) * per] == r + 1).all()) for r in range(world))
    msg += f"; multicast: {mul:.2f} ms, {per * 4 / mul / 1e6:.0f} GB/s out, {per * 4 * (world - 1) / mul / 1e6:.0f} GB/s in, data {'OK' if ok else 'WRONG'}"
print(msg, flush=True)
dist.barrier()

# the same copy kernel into cudaMalloc + CUDA-IPC buffers (what PeerBuffers used first), small and large
from nsparse_b200.multi_gpu import PeerBuffers  # noqa: E402

for nn in (1 << 28, 1 << 32):
    pb = PeerBuffers(ctx, pieces=0, fused=False)
    pb.ensure(nn, 16, torch.float32, dev)
    per2 = nn // world
    src2 = torch.full((per2,), rank + 1, dtype=torch.int32, device=dev)
    ms = timed(lambda: pb.push("col", src2, rank * per2), reps=3)
    print(f"rank {rank}: IPC buffers of {nn * 4 / 2**30:.0f} GiB: {ms:.2f} ms, {per2 * 4 * (world - 1) / ms / 1e6:.0f} GB/s out", flush=True)
    # symmetric memory of the same size
    del src2
    pb.release()
    torch.cuda.empty_cache()
t2 = symm_mem.empty(1 << 32, dtype=torch.int32, device=dev)
h2 = symm_mem.rendezvous(t2, dist.group.WORLD.group_name)
p2 = [int(p) for r, p in enumerate(h2.buffer_ptrs) if r != rank]
a2 = (C.c_void_p * len(p2))(*p2)
per2 = (1 << 32) // world
src2 = torch.full((per2,), rank + 1, dtype=torch.int32, device=dev)
ms = timed(lambda: ctx.check(ctx.lib.nsp_push_to_peers(ctx.handle, len(p2), a2, rank * per2 * 4, C.c_void_p(src2.data_ptr()), per2 * 4)), reps=3)
print(f"rank {rank}: symmetric memory of 16 GiB: {ms:.2f} ms, {per2 * 4 * (world - 1) / ms / 1e6:.0f} GB/s out", flush=True)
dist.barrier()
dist.destroy_process_group()
