#!/bin/bash
# gpurun -- bash scripts/gpu_sanitizer.sh : compute-sanitizer memcheck / racecheck / synccheck over
#   1. the reference's unchanged drivers on data/test.mtx (SpGEMM + AMB conversion + AMB SpMV),
#   2. a 64 x 4096 times 4096 x 200000 product with every row class in use (default thresholds),
#   3. the same with the class thresholds forced down and a 2^16-column window / 128-entry chunks, so that the
#      heavy kernels run windows x chunks x slabs (one row of A has 1500 entries) in both staging modes.
# Logs go to gpurun_out/sanitizer/ (copied to profiles/ by hand).
mkdir -p gpurun_out/sanitizer
O=gpurun_out/sanitizer
python scripts/make_sanitize_inputs.py /tmp > $O/inputs.txt 2>&1
CS="compute-sanitizer --print-limit 20"
FORCE="sym_bitmap_min=64,num_bitmap_min=64,sym_window_shift=16,num_window_shift=16,num_cap=128"
for tool in memcheck racecheck synccheck; do
  for p in d s; do
    timeout 900 $CS --tool $tool bin/spgemm_hash_$p tests/golden/test.mtx > $O/${tool}_spgemm_hash_${p}_testmtx.log 2>&1
    timeout 900 $CS --tool $tool bin/amb_$p tests/golden/test.mtx 2 3 > $O/${tool}_amb_${p}_testmtx.log 2>&1
    timeout 900 $CS --tool $tool oracle/_ref/dump_spgemm_ours_$p /tmp/sanitize_a_$p.bin /tmp/sanitize_b_$p.bin /tmp/san_default_$p.bin 0 > $O/${tool}_spgemm_wide_${p}_default.log 2>&1
    NSP_OPTIONS=$FORCE timeout 900 $CS --tool $tool oracle/_ref/dump_spgemm_ours_$p /tmp/sanitize_a_$p.bin /tmp/sanitize_b_$p.bin /tmp/san_forced_$p.bin 0 > $O/${tool}_spgemm_wide_${p}_forced_seg.log 2>&1
    NSP_OPTIONS=$FORCE,no_seg=1 timeout 900 $CS --tool $tool oracle/_ref/dump_spgemm_ours_$p /tmp/sanitize_a_$p.bin /tmp/sanitize_b_$p.bin /tmp/san_forced_ns_$p.bin 0 > $O/${tool}_spgemm_wide_${p}_forced_search.log 2>&1
  done
done
# the products computed under the sanitizer are the right ones
python - <<'PY' > $O/results_check.txt 2>&1
import numpy as np, sys, os
sys.path.insert(0, os.getcwd())
from oracle import oracle, refgpu
for p in "ds":
    M, K, rpt, col, val = refgpu.read_csrbin(f"/tmp/sanitize_a_{p}.bin")
    _, N, brpt, bcol, bval = refgpu.read_csrbin(f"/tmp/sanitize_b_{p}.bin")
    want = oracle.spgemm(rpt, col, val, brpt, bcol, bval, acc_double=True, n_cols=N)
    for tag in ("default", "forced", "forced_ns"):
        f = f"/tmp/san_{tag}_{p}.bin"
        if not os.path.exists(f):
            print(p, tag, "missing")
            continue
        _, _, r2, c2, v2 = refgpu.read_csrbin(f)
        print(p, tag, "OK" if np.array_equal(r2, want[0]) and np.array_equal(c2, want[1]) and np.array_equal(v2, want[2]) else "MISMATCH")
PY
for f in $O/*.log; do echo "== $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Error:|Invalid" $f | sort | uniq -c | head -8; done > $O/summary.txt
cat $O/summary.txt; cat $O/results_check.txt
