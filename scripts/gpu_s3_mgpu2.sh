mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/check_mgpu.py > gpurun_out/s3_check_mgpu.txt 2>&1; echo "check rc=$?"; grep -c OK gpurun_out/s3_check_mgpu.txt; grep -i "mismatch\|error" gpurun_out/s3_check_mgpu.txt | head -5
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e > gpurun_out/s3_bench_g2.json 2> gpurun_out/s3_bench_g2.err; echo "bench2 rc=$?"
cat gpurun_out/s3_bench_g2.json | cut -c1-200
