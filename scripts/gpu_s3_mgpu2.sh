mkdir -p gpurun_out
for extra in "" "--ip-partition" "--gather push"; do
tag=$(echo $extra | tr -d ' -')
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e $extra > gpurun_out/s3_bench_g2_$tag.json 2> gpurun_out/s3_bench_g2.err; echo "bench2 $extra rc=$?"
tail -2 gpurun_out/s3_bench_g2.err | cut -c1-300
python -c "
import json,sys
d=json.loads(open('gpurun_out/s3_bench_g2_$tag.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['config'].get('per_rank'))"
done
