#!/bin/bash
# gpurun --gpus N -- bash scripts/gpu_mgpu_configs.sh N [full] : bench.py at N GPUs on C2 and on C4 / C5 (reduced sizes
# unless "full" is given), one JSON line each
N=$1; FULL=$2
mkdir -p gpurun_out
run() {
  tag=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-e2e "$@" > gpurun_out/r2_bench_g${N}_$tag.json 2> gpurun_out/r2_bench_g${N}_$tag.err; echo "bench N=$N $tag rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_g${N}_$tag.json').read().strip().splitlines()[-1])
    g=d.get('gather') or {}
    print('$tag', round(d['ms_per_step'],2), 'ms', round(d['value'],1), 'GFLOPS nnzC', d['config']['nnz_C'], 'IP', d['config']['intermediate_products'], 'gather_ok', g.get('gather_ok'), 'no_gather', g.get('ms_no_gather'), 'parity', (d.get('parity') or {}).get('ok'), (d.get('parity') or {}).get('val_max_rel'), d['config'].get('per_rank',{}).get('nnz_c'))
except Exception as e:
    print('no line', e)
PY
  grep -E "Error|error" gpurun_out/r2_bench_g${N}_$tag.err | tail -3 | cut -c1-300
}
if [ "$FULL" = "full" ]; then
  run c2 --steps 5 --warmup 3 --cpu-seconds 6
  run c4 --config c4 --steps 3 --warmup 2
  run c5 --config c5 --steps 3 --warmup 2
  # A/B of the SM-store tail of the gather (csrc/peer_dma.cu): copy engines only, same partition model
  run c2_ce_only --steps 3 --warmup 3 --no-cpu --no-check --gather-sm 0
else
  run c2 --steps 3 --warmup 3 --cpu-seconds 5
  run c4small --config c4 --scale 19 --ef 32 --steps 2 --warmup 2
  run c5small --config c5 --c5-rows 1048576 --steps 2 --warmup 2
fi
