#!/bin/bash
# gpurun -- bash scripts/gpu_hash_ranges.sh : tests of num_hash_ranges_kernel, then configs C5 and C4 at full size on one GPU
mkdir -p gpurun_out
timeout 80 python -m pytest tests/test_spgemm_gpu.py tests/test_configs_gpu.py -m gpu -x -q -k "hash_ranges or ranges or c5_reduced" > gpurun_out/r2_pytest_hash_ranges.txt 2>&1; tail -3 gpurun_out/r2_pytest_hash_ranges.txt
run() {
  tag=$1; shift
  timeout 50 python bench.py --no-cpu --no-e2e --no-spmv --no-ref-gpu --steps 2 --warmup 1 "$@" > gpurun_out/r2_hash_ranges_$tag.json 2> gpurun_out/r2_hash_ranges_$tag.err; echo "$tag rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_hash_ranges_$tag.json').read().strip().splitlines()[-1])
    print('$tag', round(d['ms_per_step'],1), 'ms', {k: round(v['ms'],1) for k,v in d['roofline']['kernels'].items()})
except Exception as e:
    print('no line', e)
PY
}
run c5 --config c5
run c4 --config c4
