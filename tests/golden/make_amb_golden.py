"""Generates tests/golden/amb_*.npz by running the REFERENCE's own convert_amb.cu / kernel_spmv_amb.cu
(oracle/_ref/dump_amb_{d,s}, built by `make -C oracle ref_gpu` in the build container) on a GPU.

    python tests/golden/make_amb_golden.py [outdir]          # on the B200 box (gpurun)

The cases avoid the two input classes on which the reference itself is not well defined
(DESIGN.md section 6): M < 32768 with M % 32 != 0 (out-of-bounds sort flag) and rows with unsorted or
duplicate columns together with block_size > 1 (blocks are counted and filled inconsistently).
Each fixture holds the CSR input, the plan, every sfAMB array and the reference's y for a fixed x.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))

# name, M, N, density, seg_size, block_size, precision
CASES = [
    ("tiny_d", 64, 90, 0.15, 4, 2, "d"),
    ("small_seg1024_bs3_d", 4096, 5000, 0.002, 1024, 3, "d"),
    ("window_seg2048_bs1_d", 40000, 3000, 0.0012, 2048, 1, "d"),
    ("tall_seg65536_bs2_s", 70016, 150000, 0.00001, 65536, 2, "s"),
    ("band_seg4096_bs5_d", 2048, 9000, 0.0, 4096, 5, "d"),
    ("padrows_seg3072_bs4_d", 33003, 6000, 0.0005, 3072, 4, "d"),
]


def make_matrix(name, M, N, dens, seed):
    rng = np.random.default_rng(seed)
    if dens == 0.0:     # banded: runs of consecutive columns so that blocks really fill
        rows, cols = [], []
        for i in range(M):
            c0 = (i * 4) % (N - 12)
            for c in list(range(c0, c0 + 7)) + [c0 + 9]:
                rows.append(i)
                cols.append(c)
        a = sp.csr_matrix((np.ones(len(rows)), (rows, cols)), shape=(M, N))
    else:
        a = sp.random(M, N, density=dens, random_state=rng, format="csr")
    a.data = np.round(rng.random(a.nnz) * 64 + 1) / 8.0      # exact in fp32, written exactly in %g
    a.sort_indices()
    return a


def write_mtx(path, a):
    a = a.tocoo()
    order = np.lexsort((a.col, a.row))
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n")
        f.write(f"{a.shape[0]} {a.shape[1]} {a.nnz}\n")
        for r, c, v in zip(a.row[order], a.col[order], a.data[order]):
            f.write(f"{r + 1} {c + 1} {v:.17g}\n")


def read_dump(path):
    raw = open(path, "rb").read()
    hdr = np.frombuffer(raw, np.int32, 10)
    M, N, nnz, pad_M, c_size, nnz_amb, bs, seg, seg_num, rb = [int(x) for x in hdr]
    real = np.float64 if rb == 8 else np.float32
    off = 40
    out = dict(M=M, N=N, pad_M=pad_M, c_size=c_size, nnz=nnz_amb, block_size=bs, seg_size=seg, seg_num=seg_num)

    def take(dt, n):
        nonlocal off
        a = np.frombuffer(raw, dt, n, off).copy()
        off += a.nbytes
        return a

    out["rpt"], out["col"], out["val"] = take(np.int32, M + 1), take(np.int32, nnz), take(real, nnz)
    out["cs"], out["cl"] = take(np.int32, c_size), take(np.uint32, c_size)
    out["sellcs_col"], out["sellcs_val"] = take(np.uint16, nnz_amb // bs), take(real, nnz_amb)
    out["s_write_permutation"] = take(np.uint16, c_size * 32)
    out["s_write_permutation_offset"] = take(np.uint16, c_size)
    out["write_permutation"] = take(np.int32, c_size * 32)
    out["x"], out["y"] = take(real, N), take(real, M)
    assert off == len(raw)
    return out


def main():
    outdir = sys.argv[1] if len(sys.argv) > 1 else HERE
    os.makedirs(outdir, exist_ok=True)
    for seed, (name, M, N, dens, seg, bs, prec) in enumerate(CASES):
        exe = os.path.join(ROOT, "oracle", "_ref", f"dump_amb_{prec}")
        a = make_matrix(name, M, N, dens, 100 + seed)
        with tempfile.TemporaryDirectory() as td:
            mtx, dump = os.path.join(td, "a.mtx"), os.path.join(td, "a.bin")
            write_mtx(mtx, a)
            subprocess.run([exe, mtx, str(seg), str(bs), dump], check=True)
            d = read_dump(dump)
        assert np.array_equal(d["rpt"], a.indptr) and np.array_equal(d["col"], a.indices)
        np.savez_compressed(os.path.join(outdir, f"amb_{name}.npz"), **d)
        print(name, "ok", {k: d[k] for k in ("M", "N", "c_size", "nnz", "seg_size", "block_size")})


if __name__ == "__main__":
    main()
