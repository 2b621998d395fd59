"""Generates tests/golden/spgemm_ref_*.npz by running the REFERENCE's own GPU SpGEMM on a B200:
oracle/_ref/dump_spgemm_cpp_{s,d} (cuda-cpp/inc/HashSpGEMM_volta.hpp:974-1010, header unmodified) and
oracle/_ref/dump_spgemm_c_{s,d} (cuda-c/src/kernel/kernel_spgemm_hash_{s,d}.cu:1035-1075, `_sync` spelling),
both built in the build container by `make -C oracle ref_spgemm`.

    python tests/golden/make_spgemm_golden.py [outdir]          # on the B200 box (gpurun)

The volta header is the canonical result (it is the reference's own port to independent thread
scheduling); the cuda-c tree, which relies on implicit warp synchrony (README.md:17-19), is run on the
same input and the manifest records whether it agrees.  Small cases store the inputs and all of C;
the R-MAT cases above scale 10 store the generator parameters, C.rpt, a SHA-256 of C.col, every
`stride`-th row of C and the per-row sums of C.val (values 1..4: every sum is exact in fp32, so the
comparison is bit-exact whatever the summation order).
"""
import hashlib
import json
import os
import sys
import tempfile

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import refgpu  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def small_cases():
    """name -> (A, B or None) as scipy CSR with sorted indices."""
    out = {}
    rng = np.random.default_rng(7)
    a = sp.random(300, 200, density=0.05, random_state=rng, format="csr")
    b = sp.random(200, 400, density=0.05, random_state=rng, format="csr")
    for m in (a, b):
        m.data = np.round(m.data * 64 + 1) / 8.0
        m.sort_indices()
    out["rand_rect"] = (a, b)
    # one dense-ish row, empty rows, an empty column range
    rows = np.r_[np.zeros(150, int), rng.integers(2, 64, 300)]
    cols = np.r_[np.arange(150), rng.integers(0, 64, 300)]
    a = sp.csr_matrix((np.ones(len(rows)), (rows, cols)), shape=(64, 150))
    a.sum_duplicates()
    a.data[:] = (np.arange(a.nnz) % 7 + 1) / 4.0
    b = sp.random(150, 5000, density=0.02, random_state=rng, format="csr")
    b.data = np.round(b.data * 32 + 1) / 4.0
    b.sort_indices()
    a.sort_indices()
    out["hub_row"] = (a, b)
    return out


# The cuda-c tree (implicit warp synchrony, README.md:17-19) livelocks on sm_100 as soon as a row reaches the
# thread-block kernels (measured: R-MAT scale 8 never returns); it is only run where every row stays in the
# pwarp class, under a short timeout.
C_TREE_CASES = {"test_mtx"}
ONLY = set(os.environ.get("NSP_GOLDEN_ONLY", "").split(",")) - {""}

RMAT_FULL = [(8, 8, "uniform"), (10, 16, "small_int")]
RMAT_SAMPLED = [(12, 16, "small_int", 37), (14, 16, "small_int", 211)]


def compare(c1, c2, tol):
    if c1 is None or c2 is None:
        return "missing"
    (M1, N1, r1, k1, v1), (M2, N2, r2, k2, v2) = c1, c2
    if (M1, N1) != (M2, N2) or not np.array_equal(r1, r2):
        return "rpt differs"
    if not np.array_equal(k1, k2):
        return "col differs"
    err = np.abs(v1.astype(np.float64) - v2) / np.maximum(np.abs(v2.astype(np.float64)), 1e-300)
    if len(err) and err.max() > tol:
        return f"val differs (max rel {err.max():.3e})"
    return "agree"


def main():
    outdir = sys.argv[1] if len(sys.argv) > 1 else HERE
    os.makedirs(outdir, exist_ok=True)
    import nsparse_b200.gen as gen

    mpath = os.path.join(outdir, "spgemm_ref_manifest.json")
    manifest = json.load(open(mpath)) if os.path.exists(mpath) else {}
    with tempfile.TemporaryDirectory() as td:
        def run_both(name, prec, a_path, b_path):
            res, outs = {}, {}
            for tree in ("cpp", "c"):
                o = os.path.join(td, f"{name}_{tree}_{prec}.bin")
                if tree == "c" and name not in C_TREE_CASES:
                    res[tree], outs[tree] = {"error": "not run: the cuda-c kernels hang on sm_100 (implicit warp synchrony)"}, None
                    continue
                r = refgpu.run_spgemm(tree, prec, a_path, b_path, o, reps=1, timeout=45 if tree == "c" else 240)
                res[tree] = r
                outs[tree] = refgpu.read_csrbin(o) if "error" not in r and os.path.exists(o) else None
            tol = 1e-6 if prec == "s" else 1e-12
            res["c_vs_cpp"] = compare(outs["c"], outs["cpp"], tol)
            print(name, prec, {t: (r if isinstance(r, str) else r.get("error", f"nnz_c={r.get('nnz_c')} {r.get('ms_mean')} ms"))
                               for t, r in res.items()}, flush=True)
            manifest[f"{name}_{prec}"] = res
            json.dump(manifest, open(os.path.join(outdir, "spgemm_ref_manifest.json"), "w"), indent=1)
            return res, outs

        def wanted(name):
            return not ONLY or name in ONLY

        # ---- data/test.mtx through the reference's own reader ----
        mtx = os.path.join(HERE, "test.mtx")
        for prec in (("d", "s") if wanted("test_mtx") else ()):
            res, outs = run_both("test_mtx", prec, mtx, None)
            manifest[f"test_mtx_{prec}"] = res
            c = outs["cpp"] or outs["c"]
            if c is None:
                continue
            np.savez_compressed(os.path.join(outdir, f"spgemm_ref_test_mtx_{prec}.npz"), kind="mtx", c_M=c[0], c_N=c[1],
                                c_rpt=c[2], c_col=c[3], c_val=c[4], source="cpp" if outs["cpp"] else "c")

        # ---- small random cases, everything stored ----
        for name, (a, b) in small_cases().items():
            if not wanted(name):
                continue
            for prec, dt in (("d", np.float64), ("s", np.float32)):
                ap, bp = os.path.join(td, f"{name}_a_{prec}.bin"), os.path.join(td, f"{name}_b_{prec}.bin")
                refgpu.write_csrbin(ap, a.shape[0], a.shape[1], a.indptr, a.indices, a.data.astype(dt))
                refgpu.write_csrbin(bp, b.shape[0], b.shape[1], b.indptr, b.indices, b.data.astype(dt))
                res, outs = run_both(name, prec, ap, bp)
                manifest[f"{name}_{prec}"] = res
                c = outs["cpp"] or outs["c"]
                if c is None:
                    continue
                np.savez_compressed(os.path.join(outdir, f"spgemm_ref_{name}_{prec}.npz"), kind="full",
                                    a_M=a.shape[0], a_N=a.shape[1], a_rpt=a.indptr, a_col=a.indices, a_val=a.data.astype(dt),
                                    b_M=b.shape[0], b_N=b.shape[1], b_rpt=b.indptr, b_col=b.indices, b_val=b.data.astype(dt),
                                    c_M=c[0], c_N=c[1], c_rpt=c[2], c_col=c[3], c_val=c[4],
                                    source="cpp" if outs["cpp"] else "c")

        # ---- R-MAT A^2, everything stored ----
        for scale, ef, values in RMAT_FULL:
            if not wanted(f"rmat_s{scale}_ef{ef}"):
                continue
            for prec, dt in (("d", np.float64), ("s", np.float32)):
                a = gen.rmat_csr(scale, ef, seed=12345, dtype=dt, values=values)
                name = f"rmat_s{scale}_ef{ef}"
                ap = os.path.join(td, f"{name}_{prec}.bin")
                refgpu.write_csrbin(ap, a.M, a.N, a.rpt, a.col, a.val)
                res, outs = run_both(name, prec, ap, None)
                manifest[f"{name}_{prec}"] = res
                c = outs["cpp"] or outs["c"]
                if c is None:
                    continue
                np.savez_compressed(os.path.join(outdir, f"spgemm_ref_{name}_{prec}.npz"), kind="full",
                                    a_M=a.M, a_N=a.N, a_rpt=a.rpt, a_col=a.col, a_val=a.val,
                                    c_M=c[0], c_N=c[1], c_rpt=c[2], c_col=c[3], c_val=c[4],
                                    gen=json.dumps(dict(scale=scale, ef=ef, seed=12345, values=values)),
                                    source="cpp" if outs["cpp"] else "c")

        # ---- R-MAT A^2, hashed + sampled ----
        for scale, ef, values, stride in RMAT_SAMPLED:
            if not wanted(f"rmat_s{scale}_ef{ef}"):
                continue
            for prec, dt in (("s", np.float32), ("d", np.float64)):
                a = gen.rmat_csr(scale, ef, seed=12345, dtype=dt, values=values)
                name = f"rmat_s{scale}_ef{ef}"
                ap = os.path.join(td, f"{name}_{prec}.bin")
                refgpu.write_csrbin(ap, a.M, a.N, a.rpt, a.col, a.val)
                res, outs = run_both(name, prec, ap, None)
                manifest[f"{name}_{prec}"] = res
                c = outs["cpp"] or outs["c"]
                if c is None:
                    continue
                M, N, rpt, col, val = c
                rows = np.arange(stride // 2, M, stride)
                lens = (rpt[rows + 1] - rpt[rows]).astype(np.int64)
                srpt = np.concatenate([[0], np.cumsum(lens)])
                idx = np.repeat(rpt[rows].astype(np.int64) - srpt[:-1], lens) + np.arange(int(srpt[-1]))
                rowsum = np.add.reduceat(np.r_[val.astype(np.float64), 0.0], np.minimum(rpt[:-1], len(val)))
                rowsum[np.diff(rpt) == 0] = 0.0
                np.savez_compressed(os.path.join(outdir, f"spgemm_ref_{name}_{prec}.npz"), kind="sampled",
                                    gen=json.dumps(dict(scale=scale, ef=ef, seed=12345, values=values)),
                                    a_sha=sha(a.rpt) + sha(a.col) + sha(a.val), c_M=M, c_N=N, c_nnz=len(col), c_rpt=rpt,
                                    c_col_sha=sha(col), sample_rows=rows, sample_rpt=srpt, sample_col=col[idx],
                                    sample_val=val[idx], rowsum=rowsum, source="cpp" if outs["cpp"] else "c")
    json.dump(manifest, open(os.path.join(outdir, "spgemm_ref_manifest.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
