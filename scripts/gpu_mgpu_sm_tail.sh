#!/bin/bash
# gpurun --gpus N -- bash scripts/gpu_mgpu_sm_tail.sh N : the copy-engine + SM-store gather (csrc/peer_dma.cu) -- its
# one-GPU tests (1 and 3 peer buffers on the same GPU), the multi-GPU parity tests, bench.py at N GPUs and its timeline
N=$1; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_spgemm_gpu.py -k tile_pusher -m gpu -x -q > gpurun_out/r2_pytest_sm_tail_$N.txt 2>&1; tail -4 gpurun_out/r2_pytest_sm_tail_$N.txt
timeout 600 python -m pytest tests/test_multi_gpu_gpu.py "tests/test_drivers_gpu.py::test_mgpu_driver_self_check" -m gpu -x -q >> gpurun_out/r2_pytest_sm_tail_$N.txt 2>&1; tail -4 gpurun_out/r2_pytest_sm_tail_$N.txt
run_bench() {
  tag=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-e2e "$@" > gpurun_out/r2_bench_g${N}_$tag.json 2> gpurun_out/r2_bench_g${N}_$tag.err; echo "bench N=$N $tag rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_g${N}_$tag.json').read().strip().splitlines()[-1])
    print('$tag', round(d['ms_per_step'],2), 'ms', round(d['value'],1), 'GFLOPS', d['config'].get('per_rank'), d.get('gather'), (d.get('parity') or {}).get('ok'))
except Exception as e:
    print('no line', e)
PY
  grep -E "nsp dma|Error|error" gpurun_out/r2_bench_g${N}_$tag.err | tail -4 | cut -c1-400
}
run_bench smtail --steps 4 --warmup 3 --cpu-seconds 4 "$@"
NSP_DMA_TRACE=1 run_bench smtail_trace --steps 1 --warmup 3 --no-cpu --no-check "$@"
