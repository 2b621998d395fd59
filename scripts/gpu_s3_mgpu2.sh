mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/check_mgpu.py > gpurun_out/s3_check_mgpu.txt 2>&1; echo "check rc=$?"; grep -c OK gpurun_out/s3_check_mgpu.txt; grep -i "mismatch\|error" gpurun_out/s3_check_mgpu.txt | head -5
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e --ip-partition > gpurun_out/s3_bench_g2.json 2> gpurun_out/s3_bench_g2.err; echo "bench2 rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/s3_bench_g2.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['config'].get('per_rank'), {k:round(v['ms'],1) for k,v in d['roofline']['kernels'].items()})"
