"""Regenerates tests/golden/*.json and tests/golden/test.mtx.

Run in the BUILD container (needs /root/reference): `python tests/golden/make_golden.py`.
  * test.mtx        -- the reference's only fixture (data/test.mtx), re-emitted entry by entry
                       (same header kind, same entries, same order) so the GPU box has it;
  * test_mtx.json   -- CSR of that file as read by the REFERENCE's own reader
                       (oracle/_ref, nsparse.cu:14-136), C = A*A and y = A*[1..5] computed with
                       SciPy/NumPy in exact integer arithmetic (SURVEY.md 8c lists the same numbers);
  * reader_cases.json -- small MatrixMarket texts (general / symmetric / pattern) with the CSR the
                       reference reader produces for them.
"""
import json
import os
import sys
import tempfile

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import oracle  # noqa: E402

REF = "/root/reference/data/test.mtx"

READER_CASES = {
    "general_3x4": "%%MatrixMarket matrix coordinate real general\n% c\n3 4 5\n1 1 1.5\n3 4 -2\n2 2 3\n1 4 4\n3 1 0.25\n",
    "symmetric_unsorted": "%%MatrixMarket matrix coordinate real symmetric\n4 4 5\n4 1 7\n2 2 2\n3 2 5\n1 1 1\n4 3 9\n",
    "pattern_symmetric": "%%MatrixMarket matrix coordinate pattern symmetric\n3 3 3\n2 1\n3 1\n3 3\n",
}


def main():
    ref = oracle.ReferenceHost(np.float64)
    # re-emit the fixture
    lines = [l.rstrip("\n") for l in open(REF)]
    body = [l for l in lines[1:] if not l.startswith("%")]
    with open(os.path.join(HERE, "test.mtx"), "w") as f:
        f.write(lines[0] + "\n%\n" + "\n".join(body))
    a = ref.read_mtx(os.path.join(HERE, "test.mtx"))
    a0 = ref.read_mtx(REF)
    for k in ("rpt", "col", "val"):
        assert np.array_equal(a[k], a0[k])
    A = sp.csr_matrix((a["val"], a["col"], a["rpt"]), shape=(a["M"], a["N"]))
    Cm = (A @ A).tocsr()
    Cm.sort_indices()
    x = np.arange(1, a["N"] + 1, dtype=np.float64)
    ip = [int(sum(a["rpt"][c + 1] - a["rpt"][c] for c in a["col"][a["rpt"][i]:a["rpt"][i + 1]])) for i in range(a["M"])]
    out = dict(M=a["M"], N=a["N"], nnz=a["nnz"], nnz_max=a["nnz_max"], rpt=a["rpt"].tolist(), col=a["col"].tolist(),
               val=a["val"].tolist(), intprod_per_row=ip, flop=2 * sum(ip),
               c_nnz=int(Cm.nnz), c_rpt=Cm.indptr.tolist(), c_col=Cm.indices.tolist(), c_val=Cm.data.tolist(),
               x=x.tolist(), y=ref.csr_kernel(a["rpt"], a["col"], a["val"], x).tolist())
    json.dump(out, open(os.path.join(HERE, "test_mtx.json"), "w"), indent=1)
    cases = {}
    for name, text in READER_CASES.items():
        with tempfile.NamedTemporaryFile("w", suffix=".mtx", delete=False) as f:
            f.write(text)
        r = ref.read_mtx(f.name)
        os.unlink(f.name)
        cases[name] = dict(text=text, M=r["M"], N=r["N"], nnz=r["nnz"], nnz_max=r["nnz_max"], rpt=r["rpt"].tolist(),
                           col=r["col"].tolist(), val=r["val"].tolist())
    json.dump(cases, open(os.path.join(HERE, "reader_cases.json"), "w"), indent=1)
    print("golden vectors written:", out["c_rpt"], out["c_col"], out["c_val"], out["y"])


if __name__ == "__main__":
    main()
