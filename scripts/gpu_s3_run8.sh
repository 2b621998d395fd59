mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_configs_gpu.py -x -q > gpurun_out/s3_pytest8.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/s3_pytest8.txt
tail -12 gpurun_out/s3_pytest8.txt
