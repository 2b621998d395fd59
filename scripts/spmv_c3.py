"""Config C3 for profiling: CSR->AMB conversion and a few AMB SpMVs of the 5-point Laplacian n^2 fp64 (ncu -k regex:amb_spmv)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nsparse_b200 as ns  # noqa: E402
from nsparse_b200 import gen  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
lap = gen.laplacian5_csr(n, dtype=np.float64)
lap.memcpy()
ctx = ns.Context(0)
x = torch.from_numpy(np.random.default_rng(2024).random(lap.N)).cuda()
amb = ns.csr2amb(lap, ctx=ctx)
y = torch.empty(lap.M, dtype=torch.float64, device="cuda")
for _ in range(4):
    ns.spmv_amb(amb, x, out=y, ctx=ctx)
torch.cuda.synchronize()
print("seg", amb.seg_size, "block", amb.block_size, "c_size", amb._c.c_size, "nnz_amb", amb._c.nnz, "nnz", lap.nnz)
