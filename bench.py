#!/usr/bin/env python
"""bench.py -- headline benchmark of nsparse-b200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scale S]

A "step" is one complete hash SpGEMM C = A*A (symbolic + numeric phase, i.e. one
spgemm_kernel_hash call of the reference, kernel_spgemm_hash_d.cu:1035-1075) on the R-MAT
scale-20 edge-factor-16 fp32 matrix of BASELINE.json configs[1].  The line also carries the AMB
SpMV of configs[2] (5-point Laplacian 4096^2, fp64) under "spmv".

  value        GFLOPS = 2 * intermediate products / time (get_spgemm_flop), inputs resident in HBM
  e2e          the same through the host-buffer C ABI entry point (nsp_spgemm_host_s): H2D of A
               from pinned memory, both phases, D2H of all of C through a pinned staging buffer
  roofline     dominant kernel: algorithmic bytes / CUDA-event time / HBM peak
  cpu_baseline the CPU oracle (oracle/oracle.c, OpenMP, all host threads) on a bounded row sample
  N > 1        A is 1-D row-blocked by equal intermediate products, B replicated, each rank runs the
               single-GPU pipeline on its block, then an NCCL allgatherv of C's row blocks (strong
               scaling: the matrix is fixed).

--impl reference times the CPU restatement of the reference algorithm (the reference has no CPU
SpGEMM and its CUDA sources do not build for sm_100a unpatched) on all host threads.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HBM_FALLBACK_GBS = 6650.0   # /opt/skills/guides/B200_PROFILING.md fallback


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            for k in ("hbm_gbs", "hbm_gb_s", "hbm_copy_gbs"):
                if k in d:
                    return float(d[k]), "measured"
        except Exception:
            pass
    return HBM_FALLBACK_GBS, "fallback"


# ---------------------------------------------------------------------------------------------
# clocks during the timed region
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.idx = gpu_index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["sm_max_mhz"] = float(max(mx))
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


# ---------------------------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------------------------
def make_rmat(scale, ef, dtype):
    from nsparse_b200 import gen

    return gen.rmat_csr(scale, ef, seed=12345, dtype=dtype)


def strided_rows(a, stride, offset=0):
    from nsparse_b200 import CSR

    rows = np.arange(offset, a.M, stride)
    lens = (a.rpt[rows + 1] - a.rpt[rows]).astype(np.int64)
    rpt = np.zeros(len(rows) + 1, np.int64)
    rpt[1:] = np.cumsum(lens)
    idx = np.repeat(a.rpt[rows].astype(np.int64) - rpt[:-1], lens) + np.arange(int(rpt[-1]), dtype=np.int64)
    return CSR(len(rows), a.N, rpt.astype(np.int32), a.col[idx], a.val[idx], f"{a.matrix_name}[::{stride}]")


def alg_bytes_spgemm(ip, nnz_a, nnz_c, m, v):
    """bytes_alg = IP*(8+V) + nnzA*(24+V) + nnzC*(4+V) + 16*M."""
    return ip * (8 + v) + nnz_a * (24 + v) + nnz_c * (4 + v) + 16 * m


def alg_bytes_kernel(name, rows, ip, alen, nout, v):
    """Per-launch share of the same model for one row-class kernel."""
    if name.startswith("sym"):
        return 4 * ip + 12 * alen + 8 * rows
    return (4 + v) * ip + (12 + v) * alen + (4 + v) * nout + 8 * rows


# ---------------------------------------------------------------------------------------------
# CPU oracle timing (cpu_baseline leg and --impl reference)
# ---------------------------------------------------------------------------------------------
def cpu_spgemm_sample(a, target_s=15.0, steps=1, warmup=0):
    """Times the CPU oracle on every `stride`-th row of A (B = A complete).  The stride is chosen
    from a short calibration run so one step takes ~target_s."""
    from oracle import oracle

    nthr = oracle.num_threads()
    blen = np.diff(a.rpt).astype(np.int64)
    total_ip = int(blen[a.col].sum())

    def run(sub):
        t = time.perf_counter()
        c = oracle.spgemm(sub.rpt, sub.col, sub.val, a.rpt, a.col, a.val, acc_double=False, n_cols=a.N)
        dt = time.perf_counter() - t
        return dt, int(c[0][-1])

    def ip_of(sub):
        return int(blen[sub.col].sum())

    stride = max(1, int(total_ip // 2e8))           # ~2e8 products for calibration
    sub = strided_rows(a, stride, offset=stride // 2)
    dt, _ = run(sub)
    rate = ip_of(sub) / max(dt, 1e-6)
    want_ip = rate * target_s
    stride = max(1, int(total_ip / max(want_ip, 1)))
    sub = strided_rows(a, stride, offset=stride // 2)
    ip = ip_of(sub)
    for _ in range(warmup):
        run(sub)
    times = []
    for _ in range(steps):
        dt, nnzc = run(sub)
        times.append(dt)
    t = float(np.mean(times))
    return {"value": 2.0 * ip / t / 1e9, "unit": "GFLOPS", "cores": nthr, "kind": "port",
            "sample": f"every {stride}th row of A times all of B: {sub.M} rows, {ip} intermediate products, "
                      f"nnz(C_sample)={nnzc}, {t:.2f} s per pass (symbolic+numeric, OpenMP {nthr} threads)",
            "seconds": t}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if "TORCHELASTIC_RUN_ID" in os.environ or int(os.environ.get("WORLD_SIZE", "1")) > 1:
        # torchrun pins OMP_NUM_THREADS=1 for its workers; the reference arm is ONE process on all host cores
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    dtype = np.float32
    a = make_rmat(args.scale, args.ef, dtype)
    r = cpu_spgemm_sample(a, target_s=args.cpu_seconds, steps=args.steps, warmup=min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": "SpGEMM GFLOPS (C=A^2)", "value": r["value"], "unit": "GFLOPS",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["seconds"] * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"R-MAT scale-{args.scale} edgefactor-{args.ef} CSR, C=A^2 fp32",
                   "note": "CPU restatement of the reference algorithm (oracle/oracle.c): the reference "
                           "has no CPU SpGEMM; each step is a bounded row sample"},
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": "GFLOPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import nsparse_b200 as ns

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; nsparse_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION", "INFO"):
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # NCCL prints its version banner on stdout otherwise
        dist.init_process_group("nccl", device_id=dev)
    dtype = np.float32
    tdt = torch.float32
    V = 4

    t0 = time.time()
    a = make_rmat(args.scale, args.ef, dtype)
    cuts, total_ip = ns.partition_rows_by_ip(a.rpt, a.col, a.rpt, world)
    r0, r1 = cuts[rank], cuts[rank + 1]
    a_loc = a if world == 1 else ns.row_block(a, r0, r1)
    gen_s = time.time() - t0

    ctx = ns.Context(local)
    a.memcpy(local)                    # B (replicated)
    if world > 1:
        a_loc.memcpy(local)
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # -------- one step ------------------------------------------------------------------------
    state = {}

    peers = None
    if world > 1 and not args.nccl_gather:
        peers = ns.PeerBuffers(ctx, fused=args.gather == "fused", pieces=0 if args.gather == "push" else args.pieces)

    def step():
        if world > 1:
            c = ns.spgemm_kernel_hash_mgpu(a_loc, a, cuts, a.M, total_ip, ctx, peers=peers)
        else:
            c = ns.spgemm_kernel_hash(a_loc, a, ctx)
        state["c"] = c
        return c

    for w in range(args.warmup):
        c = step()
        state.pop("c", None)
        if w == 0 and world > 1 and not args.ip_partition:
            # Re-cut the row blocks with the counts of the first product: a rank's time is its compute
            # (~ intermediate products) plus what it sends (~ its entries of C times the peers), see
            # partition_rows_by_cost.  8 bytes per entry and peer at ~700 GB/s against ~10 ps per product.
            c_rpt = c.d_rpt64.cpu().numpy()
            del c
            weight = args.nnz_weight if args.nnz_weight >= 0 else 1.15 * (world - 1)
            cuts, _ = ns.partition_rows_by_cost(a.rpt, a.col, a.rpt, c_rpt, world, weight)
            a_loc = ns.row_block(a, cuts[rank], cuts[rank + 1])
            a_loc.memcpy(local)
        c = None
    barrier()
    ctx.profile(True)
    ctx.profile_dump()
    l0 = ctx.launches
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    c = None
    for _ in range(args.steps):
        c = None                       # the previous product (78 GB at scale 20) goes before the next is made
        c = step()
        state.pop("c", None)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop() if rank == 0 else None
    launches = ctx.launches - l0
    prof = ctx.profile_dump()
    ctx.profile(False)
    nnz_c, ip = c.nnz, total_ip
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        ln = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(ln)
        launches = int(ln.item())
        mine = torch.tensor([sum(p[1] for p in prof) / args.steps, float(cuts[rank + 1] - cuts[rank])],
                            dtype=torch.float64, device=dev)
        allr = [torch.zeros(2, dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"kernel_ms": [round(float(x[0]), 2) for x in allr], "rows": [int(x[1]) for x in allr]}
    else:
        per_rank = None
    ms_step = ms / args.steps
    gflops = 2.0 * ip / ms_step / 1e6

    # -------- roofline of the dominant kernel (rank 0's launches) ----------------------------------
    agg = {}
    for name, kms, rows, kip, alen, nout in prof:
        if name.endswith("_long"):
            # side-stream launch over the SAME row class (its rows with more than 1024 entries of A); the
            # library reports the span of both launches as the time of the class itself
            continue
        d = agg.setdefault(name, {"ms": 0.0, "n": 0, "bytes": 0})
        d["ms"] += kms
        d["n"] += 1
        d["bytes"] += alg_bytes_kernel(name, rows, kip, alen, nout, V)
    peak, peak_src = hbm_peak()
    roof = None
    kernels = {}
    if agg:
        for name, d in agg.items():
            kernels[name] = {"ms": d["ms"] / d["n"], "alg_GBs": d["bytes"] / d["n"] / (d["ms"] / d["n"]) / 1e6}
        top = max(agg, key=lambda k: agg[k]["ms"])
        d = agg[top]
        ach = d["bytes"] / d["n"] / (d["ms"] / d["n"]) / 1e6      # GB/s
        traffic = None
        tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get(f"scale{args.scale}", {}).get(top)
            except Exception:
                traffic = None
        roof = {"bound": "hbm", "kernel": top, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic, "peak_source": peak_src, "alg_bytes_per_launch": d["bytes"] / d["n"],
                "ms_per_launch": d["ms"] / d["n"], "share_of_step": d["ms"] / ms,
                "whole_step": {"alg_bytes": alg_bytes_spgemm(ip, a.nnz, nnz_c, a.M, V),
                               "achieved": alg_bytes_spgemm(ip, a.nnz, nnz_c, a.M, V) / ms_step / 1e6 / max(world, 1),
                               "frac": alg_bytes_spgemm(ip, a.nnz, nnz_c, a.M, V) / ms_step / 1e6 / max(world, 1) / peak},
                "kernels": kernels}

    # -------- end to end through the host-buffer C ABI ------------------------------------------------
    e2e = None
    if not args.no_e2e:
        del c
        state.clear()
        if peers is not None:
            peers.release()                # the gathered copies of C (78 GB per rank at scale 20) are not needed any more
        torch.cuda.empty_cache()
        e2e = run_e2e(args, ctx, a, a_loc, world, rank, dev, ip)

    # -------- CPU baseline (rank 0, N = 1) ----------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        r = cpu_spgemm_sample(a, target_s=args.cpu_seconds)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}

    spmv = None
    if rank == 0 and world == 1 and not args.no_spmv and hasattr(ns, "csr2amb"):
        del a_loc
        a.release()
        torch.cuda.empty_cache()
        spmv = run_spmv(args, ctx, peak, peak_src)

    if rank == 0:
        line = {
            "metric": "SpGEMM GFLOPS (C=A^2)", "value": gflops, "unit": "GFLOPS", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"R-MAT scale-{args.scale} edgefactor-{args.ef} CSR, C=A^2 fp32",
                       "M": a.M, "nnz_A": a.nnz, "intermediate_products": ip, "nnz_C": nnz_c,
                       "generator": "Graph500 Kronecker (.57,.19,.19,.05), seed 12345, no permutation, "
                                    "duplicates merged", "gen_seconds": gen_s,
                       "l2_policy": "each step writes nnz_C*8 bytes of C (>> 126 MB L2) and streams A/B "
                                    "(132 MB); no explicit flush",
                       "parallelism": (f"row-block x{world} by equal intermediate products, B replicated, allgatherv of C: "
                                       + ("NCCL broadcasts" if args.nccl_gather else
                                          "a copy kernel stores the finished block into all peers over NVLink "
                                          "(nsp_push_to_peers)" if args.gather == "push" else
                                          "fused: the numeric kernels store every chunk of C into all peers over NVLink "
                                          "as they produce it (nsp_spgemm_set_peers)" if args.gather == "fused" else
                                          f"pipelined: the block is computed in {args.pieces} pieces, the copy engines carry "
                                          "every finished piece to all peers over NVLink while the next is computed"))
                       if world > 1 else "single GPU"},
            "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clk,
        }
        if per_rank is not None:
            line["config"]["per_rank"] = per_rank
            line["config"]["partition"] = ("equal intermediate products" if args.ip_partition else
                                           "intermediate products + w * nnz(C_i), counts from the first warm-up product")
        if spmv is not None:
            line["spmv"] = spmv
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_e2e(args, ctx, a, a_loc, world, rank, dev, ip):
    """Host CSR in pinned memory -> nsp_spgemm_host_s (H2D + symbolic + numeric) -> D2H of C."""
    import torch
    import torch.distributed as dist

    L = ctx.lib
    pin = lambda x: torch.from_numpy(x).pin_memory()
    ha = [pin(a_loc.rpt), pin(a_loc.col), pin(a_loc.val)]
    same = world == 1
    hb = ha if same else [pin(a.rpt), pin(a.col), pin(a.val)]
    stage = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
    p = lambda t: C.c_void_p(t.data_ptr())
    nnz = C.c_longlong()
    csum, nbytes = C.c_ulonglong(), C.c_longlong()

    def one():
        if args.e2e_pieces > 0:
            ctx.check(L.nsp_spgemm_host_stream_s(ctx.handle, a_loc.M, a.M, a.N, p(ha[0]), p(ha[1]), p(ha[2]), p(hb[0]),
                                                 p(hb[1]), p(hb[2]), p(stage), stage.numel(), args.e2e_pieces,
                                                 C.byref(nnz), C.byref(csum), C.byref(nbytes)))
        else:
            ctx.check(L.nsp_spgemm_host_s(ctx.handle, a_loc.M, a.M, a.N, p(ha[0]), p(ha[1]), p(ha[2]), p(hb[0]),
                                          p(hb[1]), p(hb[2]), C.byref(nnz)))
            ctx.check(L.nsp_spgemm_host_drain(ctx.handle, p(stage), stage.numel(), C.byref(csum), C.byref(nbytes)))

    steps = max(1, min(args.steps, args.e2e_steps))
    one()                                   # warm-up (allocates the device buffers)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        one()
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    ms = max(e0.elapsed_time(e1), wall)     # the drain waits on the host, so take the wall clock too
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms /= steps
    h2d = sum(int(t.numel() * t.element_size()) for t in (ha if same else ha + hb))
    ctx.check(L.nsp_spgemm_host_release(ctx.handle))
    return {"value": 2.0 * ip / ms / 1e6, "unit": "GFLOPS", "h2d_bytes_per_step": h2d,
            "d2h_bytes_per_step": int(nbytes.value), "ms_per_step": ms, "steps": steps,
            "api": (f"nsp_spgemm_host_stream_s (pinned host CSR in, all of C out, {args.e2e_pieces} row ranges drained "
                    "while the next is computed)") if args.e2e_pieces > 0 else
                   "nsp_spgemm_host_s + nsp_spgemm_host_drain (pinned host CSR in, all of C out)"}


def run_spmv(args, ctx, peak, peak_src):
    """Config C3: AMB SpMV, 5-point Laplacian 4096^2 fp64 (sf_csr2amb + sf_spmv_amb)."""
    import torch

    import nsparse_b200 as ns
    from nsparse_b200 import gen

    n = args.spmv_grid
    lap = gen.laplacian5_csr(n, dtype=np.float64)
    lap.memcpy()
    x = torch.from_numpy(np.random.default_rng(2024).random(lap.N)).cuda()
    t0 = time.perf_counter()
    amb = ns.csr2amb(lap, ctx=ctx)
    torch.cuda.synchronize()
    conv_first_s = time.perf_counter() - t0      # includes lazy kernel loading and the first growth of the arena
    del amb
    t0 = time.perf_counter()
    amb = ns.csr2amb(lap, ctx=ctx)
    torch.cuda.synchronize()
    conv_s = time.perf_counter() - t0
    y = torch.empty(lap.M, dtype=torch.float64, device="cuda")
    for _ in range(5):
        ns.spmv_amb(amb, x, out=y, ctx=ctx)
    reps = 100
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    l0 = ctx.launches
    e0.record()
    for _ in range(reps):
        ns.spmv_amb(amb, x, out=y, ctx=ctx)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    alg = lap.nnz * 12 + 4 * (lap.M + 1) + 8 * (lap.N + lap.M)
    return {"metric": "AMB SpMV GFLOPS", "value": 2.0 * lap.nnz / ms / 1e6, "unit": "GFLOPS", "ms": ms,
            "GBs_alg": alg / ms / 1e6, "roofline_frac": alg / ms / 1e6 / peak, "peak_source": peak_src,
            "workload": f"5-pt Laplacian {n}^2 fp64, nnz={lap.nnz}", "seg_size": amb.seg_size,
            "block_size": amb.block_size, "conversion_s": conv_s, "conversion_first_call_s": conv_first_s, "launches": ctx.launches - l0,
            "l2_policy": "matrix values (>= 670 MB) >> L2"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=int, default=20)
    ap.add_argument("--ef", type=int, default=16)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--e2e-pieces", type=int, default=0, help="> 0: nsp_spgemm_host_stream_s with that many row ranges (measured: no gain, the D2H is PCIe-bound)")
    ap.add_argument("--spmv-grid", type=int, default=4096)
    ap.add_argument("--gather", default="fused", choices=["pipelined", "fused", "push"],
                    help="N > 1: how a rank's block of C reaches the peers (see nsparse_b200/multi_gpu.py)")
    ap.add_argument("--pieces", type=int, default=4)
    ap.add_argument("--ip-partition", action="store_true", help="N > 1: keep the equal-intermediate-products row blocks")
    ap.add_argument("--nnz-weight", type=float, default=-1.0, help="N > 1: weight of nnz(C_i) in the row cost (default 1.15 (N-1))")
    ap.add_argument("--nccl-gather", action="store_true", help="N > 1: gather C with NCCL broadcasts instead of peer stores")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-spmv", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
