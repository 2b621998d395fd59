N=$1
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 --no-e2e > gpurun_out/s3_bench_g$N.json 2> gpurun_out/s3_bench_g$N.err; echo "bench$N rc=$?"
cat gpurun_out/s3_bench_g$N.json | cut -c1-220; tail -3 gpurun_out/s3_bench_g$N.err
