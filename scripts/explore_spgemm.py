"""Development probe: per-kernel timing of the SpGEMM on an R-MAT A^2 (not the bench)."""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nsparse_b200 as ns  # noqa: E402
from nsparse_b200 import gen  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=int, default=18)
    ap.add_argument("--ef", type=int, default=16)
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--opt", action="append", default=[], help="name=value")
    ap.add_argument("--b", default="self", help="self | er4")
    ap.add_argument("--skip-check", action="store_true")
    ap.add_argument("--phases", action="store_true", help="print the cycles per phase of the heavy numeric kernel")
    ap.add_argument("--dist", action="store_true", help="print the distribution of rows / entries / products over log2(nnz(C_i))")
    ap.add_argument("--sweep", default="", help="';'-separated option sets 'k=v,k=v' run one after another")
    args = ap.parse_args()
    dt = np.float32 if args.dtype == "f32" else np.float64
    t = time.time()
    a = gen.rmat_csr(args.scale, args.ef, dtype=dt)
    b = a if args.b == "self" else gen.er_csr(a.N, a.N, 4, dtype=dt)
    print(f"gen {time.time() - t:.1f}s  M={a.M} nnzA={a.nnz} nnzB={b.nnz}", flush=True)
    ctx = ns.Context(0)
    for o in args.opt:
        k, v = o.split("=")
        ctx.set_option(k, int(v))
    a.memcpy()
    if b is not a:
        b.memcpy()
    ctx.profile(True)
    sweeps = [x for x in args.sweep.split(";")] if args.sweep else [""]
    for sw in sweeps:
      if sw:
        print(f"--- options: {sw}", flush=True)
        for o in sw.split(","):
            k, v = o.split("=")
            ctx.set_option(k, int(v))
      if args.phases:
          ctx.set_option("phase_timing", 2)
          ctx.set_option("phase_timing", 1)
      for step in range(args.steps):
          e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
          e0.record()
          d_rpt64, nnz, ip = ns.spgemm_symbolic(a, b, ctx)
          e1.record()
          d_col, d_val = ns.spgemm_numeric(a, b, d_rpt64, nnz, ctx)
          e2.record()
          torch.cuda.synchronize()
          ts, tn = e0.elapsed_time(e1), e1.elapsed_time(e2)
          prof = ctx.profile_dump()
          print(f"step {step}: symbolic {ts:.2f} ms numeric {tn:.2f} ms total {ts + tn:.2f} ms  "
                f"IP={ip} nnzC={nnz}  GFLOPS={2 * ip / (ts + tn) / 1e6:.1f}", flush=True)
          if step == args.steps - 1:
              if args.phases:
                  ctx.set_option("phase_timing", 2)
              for n, ms, rows, kip, alen, _ in prof:
                  print(f"   {n:20s} {ms:10.3f} ms rows={rows:9d} ip={kip:13d} alen={alen:11d} "
                        f"avgB={kip / max(alen, 1):8.1f}  Gprod/s={kip / max(ms, 1e-9) / 1e6:8.2f}")
          del d_col, d_val, d_rpt64
    if args.dist:
        d_rpt64, nnz, ip = ns.spgemm_symbolic(a, b, ctx)
        cnt = (d_rpt64[1:] - d_rpt64[:-1])
        blen = (b.d_rpt[1:] - b.d_rpt[:-1]).long()
        arow = torch.repeat_interleave(torch.arange(a.M, device="cuda"), (a.d_rpt[1:] - a.d_rpt[:-1]).long())
        rip = torch.zeros(a.M, dtype=torch.int64, device="cuda").index_add_(0, arow, blen[a.d_col.long()])
        alen = (a.d_rpt[1:] - a.d_rpt[:-1]).long()
        lb = torch.where(cnt > 0, torch.ceil(torch.log2(cnt.double().clamp_min(1))).long(), torch.full_like(cnt, -1))
        print("log2(nnz) bin:      rows      nnz(C) share   products share   avg E   avg products/row  avg nnz/row  compression")
        for bin_ in range(-1, 22):
            m = lb == bin_
            n = int(m.sum())
            if n == 0:
                continue
            print(f"  <=2^{bin_:2d}: {n:9d}  {float(cnt[m].sum()) / max(nnz, 1):10.4f}  {float(rip[m].sum()) / max(ip, 1):12.4f}  "
                  f"{float(alen[m].double().mean()):8.1f}  {float(rip[m].double().mean()):12.0f}  {float(cnt[m].double().mean()):10.0f}  "
                  f"{float(rip[m].sum()) / max(float(cnt[m].sum()), 1):6.2f}")
        del d_rpt64
    if args.skip_check:
        return
    # linearity check: C*1 == A*(B*1) in fp64
    d_rpt64, nnz, ip = ns.spgemm_symbolic(a, b, ctx)
    d_col, d_val = ns.spgemm_numeric(a, b, d_rpt64, nnz, ctx)
    torch.cuda.synchronize()
    rows = torch.repeat_interleave(torch.arange(a.M, device="cuda"), (d_rpt64[1:] - d_rpt64[:-1]))
    c1 = torch.zeros(a.M, dtype=torch.float64, device="cuda").index_add_(0, rows, d_val[:nnz].double())
    del rows
    brow = torch.repeat_interleave(torch.arange(b.M, device="cuda"), (b.d_rpt[1:] - b.d_rpt[:-1]).long())
    b1 = torch.zeros(b.M, dtype=torch.float64, device="cuda").index_add_(0, brow, b.d_val.double())
    arow = torch.repeat_interleave(torch.arange(a.M, device="cuda"), (a.d_rpt[1:] - a.d_rpt[:-1]).long())
    ab1 = torch.zeros(a.M, dtype=torch.float64, device="cuda").index_add_(0, arow, a.d_val.double() * b1[a.d_col.long()])
    rel = ((c1 - ab1).abs() / ab1.abs().clamp_min(1e-30)).max().item()
    print(f"row-sum check max rel err = {rel:.3e}")


if __name__ == "__main__":
    main()
