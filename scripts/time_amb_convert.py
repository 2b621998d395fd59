"""Time sf_csr2amb (nsp_csr2amb_d) on the 4096^2 Laplacian in a fresh process, several calls."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nsparse_b200 as ns
from nsparse_b200 import gen
ctx = ns.Context(0)
lap = gen.laplacian5_csr(4096, dtype=np.float64); lap.memcpy()
for i in range(4):
    torch.cuda.synchronize(); t = time.perf_counter()
    amb = ns.csr2amb(lap, ctx=ctx); torch.cuda.synchronize()
    t1 = time.perf_counter() - t
    t = time.perf_counter(); del amb; torch.cuda.synchronize(); t2 = time.perf_counter() - t
    print(f"call {i}: csr2amb {t1*1e3:.1f} ms, release {t2*1e3:.1f} ms", flush=True)
