mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/s3_smi.txt
./bin/atomics_bench > gpurun_out/s3_atomics.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s3_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/s3_pytest.txt
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/s3_bench.json 2> gpurun_out/s3_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/s3_pytest.txt; cat gpurun_out/s3_atomics.txt
