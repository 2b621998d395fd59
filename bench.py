#!/usr/bin/env python
"""bench.py -- headline benchmark of nsparse-b200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2|c4|c5] [--scale S]

A "step" is one complete hash SpGEMM C = A*B (symbolic + numeric phase, i.e. one spgemm_kernel_hash call of
the reference, kernel_spgemm_hash_d.cu:1035-1075).  Default workload: config C2 of BASELINE.json, the R-MAT
scale-20 edge-factor-16 fp32 matrix, C = A^2.  --config c4 / c5 are the 8-GPU configurations (R-MAT scale 23
ef 32 fp64 times an Erdos-Renyi B; power-law n = 16 M times the same B family).  The N = 1 line
also carries the AMB SpMV of C3 (5-point Laplacian 4096^2, fp64) under "spmv".

  value        GFLOPS = 2 * intermediate products / time (get_spgemm_flop), inputs resident in HBM
  e2e          the same through the host-buffer C ABI entry point (nsp_spgemm_host_s): H2D of A from pinned
               memory, both phases, D2H of all of C through a pinned staging buffer
  roofline     dominant kernel: algorithmic bytes / CUDA-event time / HBM peak
  parity       rows of the GPU's C against the CPU oracle's rows on the same sample (outside the timed region)
  cpu_baseline the CPU oracle (oracle/oracle.c, OpenMP, all host threads) on a bounded row sample
  ref_gpu      the REFERENCE's own GPU code (oracle/_ref, built from /root/reference in the build container) on
               the largest inputs it survives on a B200, next to this library on the same inputs
  N > 1        A is 1-D row-blocked, B replicated, each rank runs the single-GPU pipeline on its block while
               its pusher kernel stores the finished tiles of C into every peer over NVLink (allgatherv
               overlapped with the numeric phase); strong scaling (the matrix is fixed).  The line reports the
               time without the gather, the exposed gather time, the NVLink volume and gather_ok: the fold of
               the gathered C is the same on every rank and equals rank 0's single-GPU product.

--impl reference times the CPU restatement of the reference algorithm (the reference has no CPU SpGEMM and its
CUDA SpGEMM only survives inputs up to R-MAT scale 15 on a B200) on all host threads.
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
# workloads
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
    sub = CSR(len(rows), a.N, rpt.astype(np.int32), a.col[idx], a.val[idx], f"{a.matrix_name}[::{stride}]")
    sub.rows = rows
    return sub


def workload(args, local):
    """-> dict(a, b, dtype, V, name, square, on_device, describe)."""
    from nsparse_b200 import gen

    if args.config == "c2":
        a = make_rmat(args.scale, args.ef, np.float32)
        return dict(a=a, b=a, dtype=np.float32, V=4, square=True, on_device=False,
                    name=f"R-MAT scale-{args.scale} edgefactor-{args.ef} CSR, C=A^2 fp32",
                    generator="Graph500 Kronecker (.57,.19,.19,.05), seed 12345, no permutation, duplicates merged "
                              "(native generator on the host)")
    if args.config == "c4":
        scale, ef = (args.scale if args.scale != 20 else 23), (args.ef if args.ef != 16 else 32)
        a = gen.rmat_csr_device(scale, ef, seed=12345, dtype=np.float64, device=local)
        b = gen.er_csr_device(a.N, a.N, 4, seed=54321, dtype=np.float64, device=local)
        return dict(a=a, b=b, dtype=np.float64, V=8, square=False, on_device=True,
                    name=f"R-MAT scale-{scale} edgefactor-{ef} fp64 times Erdos-Renyi B (K=N=2^{scale}, 4 nnz/row), C=A*B",
                    generator="A: Graph500 Kronecker as C2 (same counter-based streams, generated on the GPU); B: 4 uniform "
                              "distinct sorted columns per row, seed 54321 (BASELINE.json does not pin B: SURVEY.md 8d)")
    if args.config == "c5":
        n = args.c5_rows
        a = gen.powerlaw_csr_device(n, 64, min(65536, n), seed=777, dtype=np.float64, device=local)
        b = gen.er_csr_device(n, n, 4, seed=54321, dtype=np.float64, device=local)
        return dict(a=a, b=b, dtype=np.float64, V=8, square=False, on_device=True,
                    name=f"power-law CSR n={n} avg 64 nnz/row max row {min(65536, n)} fp64 times Erdos-Renyi B (4 nnz/row), C=A*B",
                    generator="A: Pareto(1.5) row lengths truncated to [1, 65536], one row of exactly 65536, one uniform column "
                              "per stratum, seed 777; B as C4 (generated on the GPU)")
    raise SystemExit(f"unknown config {args.config}")


def alg_bytes_spgemm(ip, nnz_a, nnz_c, m, v):
    """bytes_alg = IP*(8+V) + nnzA*(24+V) + nnzC*(4+V) + 16*M."""
    return ip * (8 + v) + nnz_a * (24 + v) + nnz_c * (4 + v) + 16 * m


def alg_bytes_kernel(name, rows, ip, alen, nout, v):
    """Per-launch share of the same model for one row-class kernel."""
    if name.startswith("sym"):
        return 4 * ip + 12 * alen + 8 * rows
    return (4 + v) * ip + (12 + v) * alen + (4 + v) * nout + 8 * rows


# ---------------------------------------------------------------------------------------------
# CPU oracle: timing (cpu_baseline leg and --impl reference) and the parity sample
# ---------------------------------------------------------------------------------------------
def cpu_spgemm_sample(a, b, target_s=15.0, steps=1, warmup=0, acc_double=False):
    """Times the CPU oracle on every `stride`-th row of A (B complete).  The stride is chosen from a short
    calibration run so one step takes ~target_s.  Returns the timing record and (sub, c) of the last pass."""
    from oracle import oracle

    nthr = oracle.num_threads()
    blen = np.diff(b.rpt).astype(np.int64)
    total_ip = int(blen[a.col].sum())

    def run(sub):
        t = time.perf_counter()
        c = oracle.spgemm(sub.rpt, sub.col, sub.val, b.rpt, b.col, b.val, acc_double=acc_double, n_cols=b.N)
        return time.perf_counter() - t, c

    def ip_of(sub):
        return int(blen[sub.col].sum())

    stride = max(1, int(total_ip // 2e8))           # ~2e8 products for calibration
    sub = strided_rows(a, stride, offset=stride // 2)
    dt, _ = run(sub)
    rate = ip_of(sub) / max(dt, 1e-6)
    stride = max(1, int(total_ip / max(rate * target_s, 1)))
    sub = strided_rows(a, stride, offset=stride // 2)
    ip = ip_of(sub)
    for _ in range(warmup):
        run(sub)
    times, c = [], None
    for _ in range(steps):
        dt, c = run(sub)
        times.append(dt)
    t = float(np.mean(times))
    rec = {"value": 2.0 * ip / t / 1e9, "unit": "GFLOPS", "cores": nthr, "kind": "port",
           "sample": f"every {stride}th row of A times all of B: {sub.M} rows, {ip} intermediate products, "
                     f"nnz(C_sample)={int(c[0][-1])}, {t:.2f} s per pass (symbolic+numeric, OpenMP {nthr} threads)",
           "seconds": t}
    return rec, sub, c


def compare_rows(c_dev, rows, oc, V):
    """Rows `rows` of the device product against the oracle's (rpt, col, val) for those rows, compared ON the
    GPU.  Structure must be exact; values are reported as max relative error and the share above the north-star
    tolerance (1e-6 fp32 / 1e-12 fp64; the reference's own comparator, nsparse.cu:300-353, allows 1e-5 / 1e-8)."""
    import torch

    dev = c_dev.d_rpt64.device
    o_rpt, o_col, o_val = oc
    rows_t = torch.as_tensor(np.asarray(rows, dtype=np.int64), device=dev)
    beg = c_dev.d_rpt64[rows_t]
    lens = c_dev.d_rpt64[rows_t + 1] - beg
    want_lens = torch.as_tensor(np.diff(np.asarray(o_rpt, dtype=np.int64)), device=dev)
    out = {"rows": int(len(rows)), "nnz": int(o_rpt[-1])}
    if lens.numel() != want_lens.numel() or not bool((lens == want_lens).all()):
        out.update(ok=False, why="row lengths differ")
        return out
    srpt = torch.zeros(len(rows) + 1, dtype=torch.int64, device=dev)
    srpt[1:] = torch.cumsum(lens, 0)
    tol = 1e-6 if V == 4 else 1e-12
    bad_col, bad_val, max_rel, n = 0, 0, 0.0, int(srpt[-1])
    step = 1 << 27
    for s in range(0, n, step):           # in slices: the index arrays of 1e9 entries are 8 GB each
        e = min(n, s + step)
        k = torch.arange(s, e, dtype=torch.int64, device=dev)
        r = torch.searchsorted(srpt, k, right=True) - 1
        idx = beg[r] + (k - srpt[r])
        gcol, gval = c_dev.d_col[idx], c_dev.d_val[idx].double()
        wcol = torch.as_tensor(o_col[s:e], device=dev)
        wval = torch.as_tensor(o_val[s:e], device=dev).double()
        bad_col += int((gcol != wcol).sum())
        rel = (gval - wval).abs() / wval.abs().clamp_min(1e-300)
        max_rel = max(max_rel, float(rel.max()) if rel.numel() else 0.0)
        bad_val += int((rel > tol).sum())
        del k, r, idx, gcol, gval, wcol, wval, rel
    ref_tol = 1e-5 if V == 4 else 1e-8
    out.update(ok=bool(bad_col == 0 and max_rel <= ref_tol), structure_exact=bad_col == 0, val_max_rel=max_rel,
               val_tol=tol, val_above_tol=bad_val, val_above_tol_frac=bad_val / max(n, 1), reference_comparator_tol=ref_tol)
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if "TORCHELASTIC_RUN_ID" in os.environ or int(os.environ.get("WORLD_SIZE", "1")) > 1:
        # torchrun pins OMP_NUM_THREADS=1 for its workers; the reference arm is ONE process on all host cores
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    a = make_rmat(args.scale, args.ef, np.float32)
    r, _, _ = cpu_spgemm_sample(a, a, target_s=args.cpu_seconds, steps=args.steps, warmup=min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": "SpGEMM GFLOPS (C=A^2)", "value": r["value"], "unit": "GFLOPS",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["seconds"] * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"R-MAT scale-{args.scale} edgefactor-{args.ef} CSR, C=A^2 fp32",
                   "note": "CPU restatement of the reference algorithm (oracle/oracle.c): the reference has no CPU "
                           "SpGEMM, and its GPU SpGEMM (oracle/_ref, timed under ref_gpu in the other arm) does not "
                           "survive this input; each step is a bounded row sample"},
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": "GFLOPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# the reference's own GPU code on this box (oracle/_ref), next to ours on the same inputs
# ---------------------------------------------------------------------------------------------
def run_ref_gpu(args, ctx):
    import torch

    import nsparse_b200 as ns
    from nsparse_b200 import gen
    from oracle import refgpu

    out = {"note": "reference binaries built from /root/reference in the build container (make -C oracle ref_spgemm ref_gpu); "
                   "SpGEMM: cuda-cpp HashSpGEMM_volta.hpp unmodified (the cuda-c kernels livelock on sm_100; above scale 15 "
                   "the header's launches fail, at scale 18 its global tables exceed 180 GB); protocol of spgemm_hash.cu:35-52 "
                   "(mean of 10 after 1 warm-up, cudaMalloc of C inside), the same driver source for both sides"}
    with tempfile.TemporaryDirectory() as td:
        if refgpu.available("spgemm_cpp", "s") and refgpu.available("spgemm_ours", "s"):
            a = gen.rmat_csr(args.ref_scale, 16, seed=12345, dtype=np.float32)
            path = os.path.join(td, "a.bin")
            refgpu.write_csrbin(path, a.M, a.N, a.rpt, a.col, a.val)
            ref = refgpu.run_spgemm("cpp", "s", path, None, None, reps=10, timeout=120)
            ours = refgpu.run_spgemm("ours", "s", path, None, None, reps=10, timeout=120)
            out["spgemm"] = {"input": f"R-MAT scale {args.ref_scale} ef 16, C=A^2 fp32", "reference": ref, "ours": ours}
            if "gflops" in ref and "gflops" in ours:
                out["spgemm"]["speedup"] = ours["gflops"] / max(ref["gflops"], 1e-9)
        else:
            out["spgemm"] = {"unavailable": "oracle/_ref/dump_spgemm_* not built"}
        if refgpu.available("amb", "d"):
            n = args.ref_grid
            lap = gen.laplacian5_csr(n, dtype=np.float64)
            path = os.path.join(td, "lap.bin")
            refgpu.write_csrbin(path, lap.M, lap.N, lap.rpt, lap.col, lap.val)
            ref = refgpu.run_amb("d", path, 65536, 1, None, reps=100, timeout=120)
            lap.memcpy()
            x = torch.from_numpy(np.random.default_rng(2024).random(lap.N)).cuda()
            amb = ns.csr2amb(lap, plan=ns.Plan().set_plan(65536, 1), ctx=ctx)
            y = torch.empty(lap.M, dtype=torch.float64, device="cuda")
            ms = time_spmv(ns, ctx, amb, x, y, 100)
            out["spmv"] = {"input": f"5-pt Laplacian {n}^2 fp64, seg 65536 block 1 (the reference's dense conversion "
                                    "temporaries overflow int at 4096^2)", "reference": ref,
                           "ours": {"ms_mean": ms, "gflops": 2.0 * lap.nnz / ms / 1e6}}
            if "gflops" in ref:
                out["spmv"]["speedup"] = out["spmv"]["ours"]["gflops"] / max(ref["gflops"], 1e-9)
            del amb, x, y
            lap.release()
        else:
            out["spmv"] = {"unavailable": "oracle/_ref/dump_amb_d not built"}
    return out


def time_spmv(ns, ctx, amb, x, y, reps):
    """Mean ms of `reps` SpMVs replayed from one CUDA graph (the calls are 25-270 us: issued one by one from
    Python the host, not the GPU, would set the pace)."""
    import torch

    for _ in range(3):
        ns.spmv_amb(amb, x, out=y, ctx=ctx)
    torch.cuda.synchronize()
    st = torch.cuda.Stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        ns.spmv_amb(amb, x, out=y, ctx=ctx)             # binds the context to this stream before the capture
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(10):
                ns.spmv_amb(amb, x, out=y, ctx=ctx)
        g.replay()
        st.synchronize()
        e0.record(st)
        for _ in range(max(1, reps // 10)):
            g.replay()
        e1.record(st)
        st.synchronize()
    ms = e0.elapsed_time(e1) / (max(1, reps // 10) * 10)
    ctx.use_torch_stream()
    return ms


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
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # NCCL prints its banner / INFO lines on stdout otherwise
        dist.init_process_group("nccl", device_id=dev)

    t0 = time.time()
    wl = workload(args, local)
    a, b, V = wl["a"], wl["b"], wl["V"]
    tdt = torch.float64 if V == 8 else torch.float32
    if wl["on_device"]:
        cuts, total_ip = ns.partition_rows_by_ip_device(a, b, world)
    else:
        cuts, total_ip = ns.partition_rows_by_ip(a.rpt, a.col, b.rpt, world)
    gen_s = time.time() - t0
    nnz_a = a.nnz

    ctx = ns.Context(local)
    if args.push_sms:
        ctx.set_option("push_sms", args.push_sms)
    if args.gather_tma:
        ctx.set_option("gather_tma", 1)
    if args.gather_sm != 1:
        ctx.set_option("gather_sm", args.gather_sm)
    if args.hash_order:
        ctx.set_option("hash_order", args.hash_order)
    if not wl["on_device"]:
        b.memcpy(local)                    # B (replicated); A is B for C2
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # -------- rank 0's single-GPU product, folded: the reference the gathered C is compared with --------------
    fold_ref = None
    if world > 1 and not args.no_check:
        if rank == 0:
            c1 = ns.spgemm_kernel_hash(a if wl["square"] else a.memcpy(local), b, ctx)
            fold_ref = c1.fold(ctx)
            del c1
            torch.cuda.empty_cache()
        barrier()

    a_loc = a if world == 1 else ns.row_block(a, cuts[rank], cuts[rank + 1])
    if world > 1 or not wl["square"]:
        a_loc.memcpy(local)
    if wl["on_device"] and world > 1:
        if wl["square"]:
            pass
        else:
            wl["a"] = a = None             # the rank keeps its block only
            torch.cuda.empty_cache()

    peers = None
    if world > 1 and not args.nccl_gather:
        peers = ns.PeerBuffers(ctx, fused=args.gather == "fused", pieces=0 if args.gather == "push" else args.pieces)

    n_rows = cuts[-1]

    def step():
        if world > 1:
            return ns.spgemm_kernel_hash_mgpu(a_loc, b, cuts, n_rows, total_ip, ctx, peers=peers)
        return ns.spgemm_kernel_hash(a_loc, b, ctx)

    if world > 1 and not args.ip_partition and not wl["on_device"] and peers is not None and peers.fused:
        # the feedback cut needs two products to settle (and the timing rules ask for >= 3 warm-up steps anyway)
        args.warmup = max(args.warmup, 3)
    for w in range(args.warmup):
        c = step()
        if w < args.warmup - 1 and world > 1 and not args.ip_partition and not wl["on_device"] and peers is not None and peers.fused:
            # Re-cut the row blocks from what the ranks needed for the product just made and from what they have to send
            # (partition_rows_minmax: a block costs max(compute, outbound transfer)):
            # legitimate for repeated products on one pattern (the benchmark's case); a one-shot call only has the
            # equal-products cut (--ip-partition).
            c_rpt = c.d_rpt64.cpu().numpy()
            del c
            mine = torch.tensor([peers.last_compute_s], dtype=torch.float64, device=dev)
            allt = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
            dist.all_gather(allt, mine)
            secs = [float(x.item()) for x in allt]
            cuts, _ = ns.partition_rows_minmax(a.rpt, a.col, b.rpt, c_rpt, cuts, secs, world, 4 + V, args.out_gbs,
                                                 args.tail_gbs if (args.gather_sm and not args.gather_tma) else None)
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
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop() if rank == 0 else None
    launches = ctx.launches - l0
    prof = ctx.profile_dump()
    ctx.profile(False)
    nnz_c, ip = c.nnz, total_ip
    per_rank = None
    tiles_stats = (0, 0)
    if world > 1 and peers is not None and peers.fused and not args.gather_tma:
        n_ce, n_sm = C.c_longlong(0), C.c_longlong(0)
        ctx.check(ctx.lib.nsp_spgemm_peers_stats(ctx.handle, C.byref(n_ce), C.byref(n_sm), None))
        tiles_stats = (n_ce.value, n_sm.value)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        ln = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(ln)
        launches = int(ln.item())
        mine = torch.tensor([sum(p[1] for p in prof) / args.steps, float(cuts[rank + 1] - cuts[rank]),
                             float(int(c.d_rpt64[cuts[rank + 1]]) - int(c.d_rpt64[cuts[rank]]))], dtype=torch.float64, device=dev)
        allr = [torch.zeros(3, dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"kernel_ms": [round(float(x[0]), 2) for x in allr], "rows": [int(x[1]) for x in allr],
                    "nnz_c": [int(x[2]) for x in allr]}
    ms_step = ms / args.steps
    gflops = 2.0 * ip / ms_step / 1e6

    # -------- parity of a row sample against the CPU oracle + CPU baseline (rank 0) ----------------------------
    cpu, parity = None, None
    if rank == 0 and not args.no_cpu:
        if world > 1:
            from oracle import oracle as _orc

            _orc.set_threads(max(1, (os.cpu_count() or 1) // 2))      # torchrun pins OMP_NUM_THREADS=1 for its workers
        if wl["on_device"]:
            # C4 / C5: a few hundred rows spread over the matrix (the full inputs never leave the GPU)
            src = a_loc if world > 1 else a
            rows = np.unique(np.linspace(0, src.M - 1, 257).astype(np.int64))
            sub = src.rows_to_host(rows)
            hb = b.to_host()
            from oracle import oracle

            tcpu = time.perf_counter()
            oc = oracle.spgemm(sub.rpt, sub.col, sub.val, hb.rpt, hb.col, hb.val, acc_double=True, n_cols=hb.N)
            tcpu = time.perf_counter() - tcpu
            sub_ip = int(np.diff(hb.rpt).astype(np.int64)[sub.col].sum())
            cpu = {"value": 2.0 * sub_ip / max(tcpu, 1e-9) / 1e9, "unit": "GFLOPS", "cores": oracle.num_threads(), "kind": "port",
                   "sample": f"{len(rows)} rows spread over {'this rank\'s block of ' if world > 1 else ''}A times all of B: "
                             f"{sub_ip} products, {tcpu:.2f} s"}
            grows = rows + (cuts[rank] if world > 1 else 0)
        else:
            r, sub, oc = cpu_spgemm_sample(a, b, target_s=args.cpu_seconds, acc_double=True)
            cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
            grows = sub.rows
        if c is not None:
            parity = compare_rows(c, grows, oc, V)
            parity["against"] = "CPU oracle (oracle/oracle.c, pinned to the reference's GPU output by tests/golden/spgemm_ref_*.npz)"
        del oc, sub
    if world > 1:
        barrier()

    # -------- N > 1: the gathered C is the same everywhere and equals the single-GPU product -------------------
    gather = None
    if world > 1:
        gather = {}
        if not args.no_check:
            f = c.fold(ctx)
            mine = torch.tensor([f[0] & 0x7fffffffffffffff, f[1] & 0x7fffffffffffffff], dtype=torch.int64, device=dev)
            sums = torch.tensor([f[2], f[3]], dtype=torch.float64, device=dev)
            allh = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
            alls = [torch.zeros(2, dtype=torch.float64, device=dev) for _ in range(world)]
            dist.all_gather(allh, mine)
            dist.all_gather(alls, sums)
            if rank == 0:
                same = all(bool((h == allh[0]).all()) for h in allh)
                ref_h = (fold_ref[0] & 0x7fffffffffffffff, fold_ref[1] & 0x7fffffffffffffff)
                eq_ref = (int(allh[0][0]), int(allh[0][1])) == ref_h
                rel = max(abs(float(s[i]) - fold_ref[2 + i]) / max(abs(fold_ref[2 + i]), 1e-300) for s in alls for i in range(2))
                gather.update(gather_ok=bool(same and eq_ref and rel < 1e-6), ranks_equal=bool(same),
                              structure_equals_single_gpu=bool(eq_ref), value_sums_max_rel=rel,
                              how="device fold (nsp_csr_fold_*) of the gathered C on every rank: 64-bit hashes of rpt and col "
                                  "exact, value sums within 1e-6, against rank 0's single-GPU product")
        # time without the gather: every rank computes its block only
        c = None
        # (the peer buffers stay allocated: the no-gather products go to fresh memory)
        torch.cuda.empty_cache()
        for _ in range(2):
            cl = ns.spgemm_kernel_hash(a_loc, b, ctx)
            del cl
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(args.steps):
            cl = ns.spgemm_kernel_hash(a_loc, b, ctx)
            del cl
        g1.record()
        barrier()
        t = torch.tensor([g0.elapsed_time(g1) / args.steps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_ng = float(t.item())
        if rank == 0:
            out_b = max(per_rank["nnz_c"]) * (4 + V) * (world - 1)
            in_b = (nnz_c - min(per_rank["nnz_c"])) * (4 + V)
            gather.update(ms_no_gather=ms_ng, ms_gather_exposed=ms_step - ms_ng,
                          nvlink_out_bytes_per_gpu_max=out_b, nvlink_in_bytes_per_gpu_max=in_b,
                          nvlink_in_GBs_over_step=in_b / ms_step / 1e6, nvlink_out_GBs_over_step=out_b / ms_step / 1e6,
                          nvlink_floor_ms=in_b / 700e9 * 1e3, tiles_by_copy_engines_rank0=tiles_stats[0],
                          tiles_by_sm_stores_rank0=tiles_stats[1],
                          note="floor = inbound bytes at the 700 GB/s one GPU was measured to receive "
                               "(profiles/r1_probe_nvlink_multicast_2gpu.txt)")

    # -------- roofline of the dominant kernel (rank 0's launches) ----------------------------------
    agg = {}
    for name, kms, rows, kip, alen, nout in prof:
        if name.endswith("_long") or name.endswith("_own"):
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
        if os.path.exists(tp) and args.config == "c2":
            try:
                traffic = json.load(open(tp)).get(f"scale{args.scale}", {}).get(top)
            except Exception:
                traffic = None
        whole = alg_bytes_spgemm(ip, nnz_a, nnz_c, n_rows, V)
        roof = {"bound": "hbm", "kernel": top, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic, "peak_source": peak_src, "alg_bytes_per_launch": d["bytes"] / d["n"],
                "ms_per_launch": d["ms"] / d["n"], "share_of_step": d["ms"] / ms,
                "whole_step": {"alg_bytes": whole, "achieved": whole / ms_step / 1e6 / max(world, 1),
                               "frac": whole / ms_step / 1e6 / max(world, 1) / peak},
                "kernels": kernels}

    c = None
    # -------- end to end through the host-buffer C ABI ------------------------------------------------
    e2e = None
    if not args.no_e2e and args.config == "c2":
        if peers is not None:
            peers.release()                # the gathered copies of C (78 GB per rank at scale 20) are not needed any more
        torch.cuda.empty_cache()
        e2e = run_e2e(args, ctx, b, a_loc, world, rank, dev, ip)
    elif peers is not None:
        peers.release()

    ref_gpu = None
    if rank == 0 and world == 1 and not args.no_ref_gpu and args.config == "c2":
        torch.cuda.empty_cache()
        ref_gpu = run_ref_gpu(args, ctx)

    spmv = None
    if rank == 0 and world == 1 and not args.no_spmv and args.config == "c2":
        del a_loc
        if hasattr(a, "release"):
            a.release()
        torch.cuda.empty_cache()
        spmv = run_spmv(args, ctx, peak, peak_src)

    if rank == 0:
        gather_txt = ("NCCL broadcasts" if args.nccl_gather else
                      "a copy kernel stores the finished block into all peers over NVLink (nsp_push_to_peers)" if args.gather == "push" else
                      "overlapped, TMA pusher: the numeric kernels count finished tiles of C and a pusher kernel on a few SMs of its "
                      "own moves them into all peers with cp.async.bulk while the rest computes" if (args.gather == "fused" and args.gather_tma) else
                      "overlapped, copy engines: the heavy rows are computed tile by tile, the numeric kernels count finished entries "
                      "per tile, and the calling host thread hands every finished tile to cudaMemcpyAsync for each peer while the "
                      "rest is computed (nsp_spgemm_set_peers)" + ("; tiles still unsent when the kernels end are stored to all peers by "
                      "an SM copy kernel next to the copy engines" if args.gather_sm == 1 else "; SM stores only, after the kernels"
                      if args.gather_sm == 2 else "") if args.gather == "fused" else
                      f"pipelined: the block is computed in {args.pieces} pieces, the copy engines carry every finished piece")
        line = {
            "metric": "SpGEMM GFLOPS (C=A^2)" if wl["square"] else "SpGEMM GFLOPS (C=A*B)", "value": gflops, "unit": "GFLOPS",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32" if V == 4 else "f64", "data": "synthetic",
            "config": {"workload": wl["name"], "M": n_rows, "nnz_A": nnz_a, "intermediate_products": ip, "nnz_C": nnz_c,
                       "generator": wl["generator"], "gen_seconds": gen_s,
                       "l2_policy": "each step writes nnz_C*(4+V) bytes of C (>> 126 MB L2) and streams A/B; no explicit flush",
                       "parallelism": (f"row-block x{world}, B replicated, allgatherv of C: " + gather_txt) if world > 1 else "single GPU"},
            "roofline": roof, "cpu_baseline": cpu, "parity": parity, "e2e": e2e, "gpu_launches": launches, "clocks": clk,
        }
        if per_rank is not None:
            line["config"]["per_rank"] = per_rank
            line["config"]["partition"] = ("equal intermediate products" if (args.ip_partition or wl["on_device"]) else
                                           "feedback: a block costs max(compute, outbound transfer) -- rows charged their "
                                           "intermediate products at the rate their block was computed at in the previous warm-up "
                                           f"product, entries of C at {args.out_gbs:.0f} GB/s to the N-1 peers while the kernels run and {args.tail_gbs:.0f} GB/s after (partition_rows_minmax; "
                                           "repeated products on one pattern -- a one-shot call has the equal-products cut)")
        if gather is not None:
            line["gather"] = gather
        if ref_gpu is not None:
            line["ref_gpu"] = ref_gpu
        if spmv is not None:
            line["spmv"] = spmv
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_e2e(args, ctx, b, a_loc, world, rank, dev, ip):
    """Host CSR in pinned memory -> nsp_spgemm_host_s (H2D + symbolic + numeric) -> D2H of C."""
    import torch
    import torch.distributed as dist

    L = ctx.lib
    pin = lambda x: torch.from_numpy(x).pin_memory()
    ha = [pin(a_loc.rpt), pin(a_loc.col), pin(a_loc.val)]
    same = world == 1
    hb = ha if same else [pin(b.rpt), pin(b.col), pin(b.val)]
    stage = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
    p = lambda t: C.c_void_p(t.data_ptr())
    nnz = C.c_longlong()
    csum, nbytes = C.c_ulonglong(), C.c_longlong()

    def one():
        if args.e2e_pieces > 0:
            ctx.check(L.nsp_spgemm_host_stream_s(ctx.handle, a_loc.M, b.M, b.N, p(ha[0]), p(ha[1]), p(ha[2]), p(hb[0]),
                                                 p(hb[1]), p(hb[2]), p(stage), stage.numel(), args.e2e_pieces,
                                                 C.byref(nnz), C.byref(csum), C.byref(nbytes)))
        else:
            ctx.check(L.nsp_spgemm_host_s(ctx.handle, a_loc.M, b.M, b.N, p(ha[0]), p(ha[1]), p(ha[2]), p(hb[0]),
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
    from oracle import oracle

    n = args.spmv_grid
    lap = gen.laplacian5_csr(n, dtype=np.float64)
    lap.memcpy()
    hx = np.random.default_rng(2024).random(lap.N)
    x = torch.from_numpy(hx).cuda()
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
    l0 = ctx.launches
    ms = time_spmv(ns, ctx, amb, x, y, 100)
    launches = ctx.launches - l0
    alg = lap.nnz * 12 + 4 * (lap.M + 1) + 8 * (lap.N + lap.M)
    # AMB footprint with the reference's own model (convert_amb.cu:785-791)
    c_size, nnz_amb, bs = int(amb._c.c_size), int(amb._c.nnz), int(amb._c.block_size)
    fp = 2 * nnz_amb // bs + 8 * nnz_amb + 8 * c_size + 2 * 32 * c_size + 2 * c_size + 2 * 8 * 32 * c_size + 2 * 8 * lap.M
    # parity at full size: y against the reference's CPU SpMV restated (csr_kernel, nsparse.cu:240-259)
    ns.spmv_amb(amb, x, out=y, ctx=ctx)
    hy = y.cpu().numpy()
    t1 = time.perf_counter()
    want = oracle.spmv_csr(lap.rpt, lap.col, lap.val, hx, parallel=False)
    cpu_serial_s = time.perf_counter() - t1
    t1 = time.perf_counter()
    oracle.spmv_csr(lap.rpt, lap.col, lap.val, hx, parallel=True)
    cpu_omp_s = time.perf_counter() - t1
    d = np.abs(hy - want)
    rel = float((d / np.maximum(np.abs(want), 1e-300)).max())
    # y_i = 4 x_i - (neighbours) cancels: |y_i| can be 1e-5 of its terms, so the tolerance is applied to the error
    # relative to sum_j |a_ij| |x_j| (componentwise backward error); the plain relative error is reported next to it
    scale = oracle.spmv_csr(lap.rpt, lap.col, np.abs(lap.val), np.abs(hx), parallel=True)
    rel_scaled = float((d / np.maximum(scale, 1e-300)).max())
    ok = rel_scaled <= 1e-12
    # end to end through the host-buffer entry point (x in, y out over PCIe every call)
    L = ctx.lib
    hxp = torch.from_numpy(hx).pin_memory()
    hyp = torch.empty(lap.M, dtype=torch.float64).pin_memory()
    fn = L.nsp_spmv_amb_host_d
    for _ in range(2):
        ctx.check(fn(ctx.handle, C.byref(amb._c), C.c_void_p(hxp.data_ptr()), C.c_void_p(hyp.data_ptr())))
    t1 = time.perf_counter()
    reps = 10
    for _ in range(reps):
        ctx.check(fn(ctx.handle, C.byref(amb._c), C.c_void_p(hxp.data_ptr()), C.c_void_p(hyp.data_ptr())))
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t1) * 1e3 / reps
    return {"metric": "AMB SpMV GFLOPS", "value": 2.0 * lap.nnz / ms / 1e6, "unit": "GFLOPS", "ms": ms,
            "GBs_alg": alg / ms / 1e6, "roofline_frac": alg / ms / 1e6 / peak, "peak_source": peak_src,
            "alg_bytes": alg, "amb_footprint_bytes": fp, "amb_footprint_GBs": fp / ms / 1e6,
            "amb_footprint_model": "convert_amb.cu:785-791: 2*nnz_amb/bs + V*nnz_amb + 8*c + 64*c + 2*c + 2*V*32*c + 2*V*M",
            "workload": f"5-pt Laplacian {n}^2 fp64, nnz={lap.nnz}", "seg_size": amb.seg_size, "block_size": amb.block_size,
            "c_size": c_size, "nnz_amb": nnz_amb,
            "conversion_s": conv_s, "conversion_first_call_s": conv_first_s, "launches": launches,
            "timing": "100 calls replayed from CUDA graphs of 10",
            "parity": {"ok": bool(ok), "max_err_over_abs_row_sum": rel_scaled, "tol": 1e-12, "max_rel_to_y": rel,
                       "against": "csr_kernel restated (oracle.c), full size; the error of y_i is measured against sum_j |a_ij||x_j| "
                                  "(y_i itself cancels to 1e-5 of its terms on this matrix)"},
            "cpu_baseline": {"serial": {"value": 2.0 * lap.nnz / cpu_serial_s / 1e9, "unit": "GFLOPS", "cores": 1, "kind": "port",
                                        "sample": "whole matrix, csr_kernel (nsparse.cu:240-259)"},
                             "openmp": {"value": 2.0 * lap.nnz / cpu_omp_s / 1e9, "unit": "GFLOPS", "cores": oracle.num_threads(),
                                        "kind": "port", "sample": "whole matrix, row-parallel csr_kernel"}},
            "e2e": {"value": 2.0 * lap.nnz / e2e_ms / 1e6, "unit": "GFLOPS", "ms": e2e_ms, "h2d_bytes_per_step": 8 * lap.N,
                    "d2h_bytes_per_step": 8 * lap.M, "api": "nsp_spmv_amb_host_d (pinned x in, y out)"},
            "l2_policy": "matrix values (>= 670 MB) >> L2"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=["c2", "c4", "c5"])
    ap.add_argument("--scale", type=int, default=20)
    ap.add_argument("--ef", type=int, default=16)
    ap.add_argument("--c5-rows", type=int, default=1 << 24)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--e2e-pieces", type=int, default=0, help="> 0: nsp_spgemm_host_stream_s with that many row ranges (measured: no gain, the D2H is PCIe-bound)")
    ap.add_argument("--spmv-grid", type=int, default=4096)
    ap.add_argument("--ref-scale", type=int, default=14, help="R-MAT scale of the reference-GPU leg (its SpGEMM survives up to 15 on a B200)")
    ap.add_argument("--ref-grid", type=int, default=2048, help="Laplacian grid of the reference-GPU SpMV leg (its conversion overflows at 4096)")
    ap.add_argument("--gather", default="fused", choices=["pipelined", "fused", "push"],
                    help="N > 1: how a rank's block of C reaches the peers (see nsparse_b200/multi_gpu.py)")
    ap.add_argument("--pieces", type=int, default=4)
    ap.add_argument("--push-sms", type=int, default=0, help="N > 1, --gather-tma: SMs of the pusher kernel (0: library default)")
    ap.add_argument("--gather-tma", action="store_true", help="N > 1: the TMA pusher kernel instead of the copy engines")
    ap.add_argument("--ip-partition", action="store_true", help="N > 1: keep the equal-intermediate-products row blocks")
    ap.add_argument("--hash-order", type=int, default=0, help="measurements: 1 = bitonic sort of the hash tables always, 2 = no shared-memory bucket ordering")
    ap.add_argument("--gather-sm", type=int, default=1, help="N > 1: 1 = tiles left when the kernels end go out by SM stores next to the copy engines, 0 = copy engines only, 2 = SM stores only")
    ap.add_argument("--out-gbs", type=float, default=400.0, help="N > 1: outbound GB/s per GPU (copy engines, while the kernels run) assumed by the row partition")
    ap.add_argument("--tail-gbs", type=float, default=650.0, help="N > 1: outbound GB/s per GPU once the kernels have ended (copy engines + SM stores) assumed by the row partition")
    ap.add_argument("--nccl-gather", action="store_true", help="N > 1: gather C with NCCL broadcasts instead of peer stores")
    ap.add_argument("--no-check", action="store_true", help="N > 1: skip the gather_ok check")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-spmv", action="store_true")
    ap.add_argument("--no-ref-gpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
