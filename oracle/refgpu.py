"""Runner for the REFERENCE's own GPU binaries in oracle/_ref/ -- TEST INFRASTRUCTURE ONLY.

oracle/_ref/dump_spgemm_{c,cpp}_{s,d} (reference SpGEMM, cuda-c `_sync`-spelled / cuda-cpp volta header
unmodified) and oracle/_ref/dump_amb_{s,d} (reference CSR->AMB + AMB SpMV) are built in the build
container by `make -C oracle ref_spgemm ref_gpu` from the reference sources where they lie and travel to
the GPU box with the snapshot.  This module writes their raw CSR inputs, runs them (each under a
timeout: the reference's global-table fallback can ask for more memory than the GPU has, and its kernels
rely on implicit warp synchrony) and parses what they print.  Used by tests/golden/make_spgemm_golden.py
and by bench.py's reference-GPU leg; the product package never imports it.
"""
from __future__ import annotations

import json
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")


def exe(kind: str, prec: str) -> str:
    """kind: 'spgemm_c' | 'spgemm_cpp' | 'amb'; prec: 's' | 'd'."""
    return os.path.join(REF_DIR, f"dump_{kind}_{prec}")


def available(kind: str, prec: str) -> bool:
    return os.access(exe(kind, prec), os.X_OK)


def write_csrbin(path: str, M: int, N: int, rpt, col, val) -> None:
    val = np.ascontiguousarray(val)
    with open(path, "wb") as f:
        np.array([M, N, len(col), val.dtype.itemsize], np.int32).tofile(f)
        np.ascontiguousarray(rpt, np.int32).tofile(f)
        np.ascontiguousarray(col, np.int32).tofile(f)
        val.tofile(f)


def read_csrbin(path: str):
    raw = np.fromfile(path, np.uint8)
    M, N, nnz, vb = [int(x) for x in raw[:16].view(np.int32)]
    real = np.float64 if vb == 8 else np.float32
    o = 16
    rpt = raw[o:o + 4 * (M + 1)].view(np.int32).copy()
    o += 4 * (M + 1)
    col = raw[o:o + 4 * nnz].view(np.int32).copy()
    o += 4 * nnz
    val = raw[o:o + vb * nnz].view(real).copy()
    assert o + vb * nnz == len(raw), "csrbin: trailing bytes"
    return M, N, rpt, col, val


def _run(cmd, timeout):
    try:
        p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=timeout, text=True)
    except subprocess.TimeoutExpired:
        return {"error": f"timeout after {timeout} s"}
    if p.returncode != 0:
        tail = (p.stderr or p.stdout or "").strip().splitlines()[-3:]
        return {"error": f"exit code {p.returncode}: " + " | ".join(tail)}
    for line in reversed(p.stdout.splitlines()):
        line = line.strip()
        if line.startswith("{"):
            try:
                return json.loads(line)
            except ValueError:
                pass
    return {"error": "no result line", "stdout": p.stdout[-400:]}


def run_spgemm(tree: str, prec: str, a_path: str, b_path: str | None = None, out_path: str | None = None,
               reps: int = 10, timeout: float = 300.0) -> dict:
    """tree: 'c' (cuda-c spgemm_kernel_hash) or 'cpp' (HashSpGEMM_volta.hpp SpGEMM_Hash)."""
    return _run([exe(f"spgemm_{tree}", prec), a_path, b_path or "-", out_path or "-", str(reps)], timeout)


def run_amb(prec: str, a_path: str, seg_size: int, block_size: int, out_path: str | None = None, reps: int = 100,
            timeout: float = 300.0) -> dict:
    return _run([exe("amb", prec), a_path, str(seg_size), str(block_size), out_path or "-", str(reps)], timeout)
