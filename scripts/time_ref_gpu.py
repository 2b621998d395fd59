"""Same-box timing of the REFERENCE's own GPU code against this library.

    python scripts/time_ref_gpu.py [--scales 14,16,17,18] [--out gpurun_out/r2_ref_gpu.json]

SpGEMM: R-MAT scale S edge factor 16, C = A^2, fp32 (and fp64 at the smallest scale), reference protocol of
spgemm_hash.cu:35-52 (mean of 10 calls after one warm-up, cudaMalloc/cudaFree of C inside the timed region)
through ONE driver source (oracle/ref_gpu/dump_spgemm.cu) linked three ways: the reference's cuda-c kernels
(`_sync` spelling), the reference's cuda-cpp volta header (unmodified), and this library's nsparse.h archive.
SpMV: 5-point Laplacian n^2 fp64, reference sf_csr2amb + sf_spmv_amb (mean of 100 after one warm-up) against
nsp_spmv_amb_d on the same matrix and plan.
"""
import argparse
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refgpu  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scales", default="14,16,17,18")
    ap.add_argument("--grids", default="1024,2048")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r2_ref_gpu.json"))
    ap.add_argument("--timeout", type=float, default=180.0)
    args = ap.parse_args()
    import torch

    import nsparse_b200 as ns
    from nsparse_b200 import gen

    res = {"gpu": torch.cuda.get_device_name(0), "spgemm": [], "spmv": []}
    with tempfile.TemporaryDirectory() as td:
        for i, scale in enumerate(int(s) for s in args.scales.split(",")):
            for prec, dt in ((("s", np.float32), ("d", np.float64)) if i == 0 else (("s", np.float32),)):
                a = gen.rmat_csr(scale, 16, seed=12345, dtype=dt)
                path = os.path.join(td, f"rmat{scale}_{prec}.bin")
                refgpu.write_csrbin(path, a.M, a.N, a.rpt, a.col, a.val)
                row = {"input": f"R-MAT scale {scale} ef 16 A^2 fp{32 if prec == 's' else 64}", "nnz_a": a.nnz}
                for tree in ("ours", "cpp", "c"):
                    row[tree] = refgpu.run_spgemm(tree, prec, path, None, None, reps=10, timeout=args.timeout)
                    print(row["input"], tree, row[tree], flush=True)
                res["spgemm"].append(row)
                os.unlink(path)
        ctx = ns.Context(0)
        for n in (int(g) for g in args.grids.split(",")):
            lap = gen.laplacian5_csr(n, dtype=np.float64)
            path = os.path.join(td, f"lap{n}.bin")
            refgpu.write_csrbin(path, lap.M, lap.N, lap.rpt, lap.col, lap.val)
            row = {"input": f"5-pt Laplacian {n}^2 fp64", "nnz": lap.nnz}
            row["ref"] = refgpu.run_amb("d", path, 65536, 1, None, reps=100, timeout=args.timeout)
            os.unlink(path)
            lap.memcpy()
            x = torch.from_numpy(np.random.default_rng(2024).random(lap.N)).cuda()
            amb = ns.csr2amb(lap, plan=ns.Plan().set_plan(65536, 1), ctx=ctx)
            y = torch.empty(lap.M, dtype=torch.float64, device="cuda")
            for _ in range(3):
                ns.spmv_amb(amb, x, out=y, ctx=ctx)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(100):
                ns.spmv_amb(amb, x, out=y, ctx=ctx)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 100
            row["ours"] = {"ms_mean": ms, "gflops": 2.0 * lap.nnz / ms / 1e6, "seg_size": amb.seg_size, "block_size": amb.block_size}
            print(row, flush=True)
            res["spmv"].append(row)
            del amb, x, y
            lap.release()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
