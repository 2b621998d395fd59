mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s3_pytest7.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/s3_pytest7.txt
tail -8 gpurun_out/s3_pytest7.txt
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/s3_bench7.json 2> gpurun_out/s3_bench7.err; echo "bench rc=$?"; tail -3 gpurun_out/s3_bench7.err
