#!/bin/bash
# gpurun --gpus N -- bash scripts/gpu_mgpu_dev.sh N : multi-GPU parity tests, the C driver, and bench.py at N GPUs
N=$1; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 600 python -m pytest tests/test_multi_gpu_gpu.py "tests/test_drivers_gpu.py::test_mgpu_driver_self_check" -m gpu -x -q > gpurun_out/r2_pytest_mgpu_$N.txt 2>&1; tail -15 gpurun_out/r2_pytest_mgpu_$N.txt
run_bench() {
  tag=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 --no-e2e "$@" > gpurun_out/r2_bench_g${N}_$tag.json 2> gpurun_out/r2_bench_g${N}_$tag.err; echo "bench N=$N $tag rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_g${N}_$tag.json').read().strip().splitlines()[-1])
    print('$tag', round(d['ms_per_step'],2), 'ms', round(d['value'],1), 'GFLOPS', d['config'].get('per_rank',{}).get('kernel_ms'), d.get('gather'), d.get('parity'))
except Exception as e:
    print('no line', e)
PY
  tail -3 gpurun_out/r2_bench_g${N}_$tag.err | cut -c1-300
}
run_bench default --cpu-seconds 5 "$@"
run_bench ip --no-cpu --no-check --ip-partition "$@"
run_bench tma32 --no-cpu --no-check --gather-tma --push-sms 32 "$@"
