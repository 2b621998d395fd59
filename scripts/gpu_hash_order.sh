#!/bin/bash
# gpurun -- bash scripts/gpu_hash_order.sh : GPU tests, then configs C5 and C4 at full size on ONE GPU with the three
# ways of ordering a hash-class row (option hash_order): per-class times in roofline.kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_hash_order.txt 2>&1; tail -3 gpurun_out/r2_pytest_hash_order.txt
run() {
  tag=$1; shift
  timeout 300 python bench.py --no-cpu --no-e2e --no-spmv --no-ref-gpu --steps 2 --warmup 1 "$@" > gpurun_out/r2_hash_order_$tag.json 2> gpurun_out/r2_hash_order_$tag.err; echo "$tag rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_hash_order_$tag.json').read().strip().splitlines()[-1])
    print('$tag', round(d['ms_per_step'],1), 'ms', {k: round(v['ms'],1) for k,v in d['roofline']['kernels'].items()})
except Exception as e:
    print('no line', e)
PY
}
run c5_default --config c5
run c5_bitonic --config c5 --hash-order 1
run c4_default --config c4
run c4_bitonic --config c4 --hash-order 1
run c5_through_c --config c5 --hash-order 2
