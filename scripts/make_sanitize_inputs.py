"""Inputs for scripts/gpu_sanitizer.sh: a product small enough for racecheck that still drives the heavy SpGEMM
kernels through several column windows (N = 200 000 > 2 * 2^16), many accumulator chunks (num_cap = 128) and the
multi-slab instantiation (one row of A with 1500 entries), next to rows of every light class."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import refgpu  # noqa: E402

out = sys.argv[1] if len(sys.argv) > 1 else "/tmp"
rng = np.random.default_rng(3)
K, N = 4096, 200000
lens = np.array([0, 1, 2, 3, 5, 9, 17, 40, 90, 300, 700, 1024, 1025, 1500] + list(rng.integers(1, 120, 50)))
a_rpt = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
a_col = np.concatenate([np.sort(rng.choice(K, n, replace=False)) for n in lens]).astype(np.int32)
b_len = rng.integers(20, 400, K)
b_rpt = np.concatenate([[0], np.cumsum(b_len)]).astype(np.int32)
b_col = np.concatenate([np.sort(rng.choice(N, n, replace=False)) for n in b_len]).astype(np.int32)
for prec, dt in (("s", np.float32), ("d", np.float64)):
    a_val = rng.integers(1, 4, len(a_col)).astype(dt)
    b_val = rng.integers(1, 4, len(b_col)).astype(dt)
    refgpu.write_csrbin(os.path.join(out, f"sanitize_a_{prec}.bin"), len(lens), K, a_rpt, a_col, a_val)
    refgpu.write_csrbin(os.path.join(out, f"sanitize_b_{prec}.bin"), K, N, b_rpt, b_col, b_val)
ip = int(b_len[a_col].sum())
print("A", len(lens), "x", K, "nnz", len(a_col), "longest row", int(lens.max()), "| B", K, "x", N, "nnz", len(b_col), "| products", ip)
