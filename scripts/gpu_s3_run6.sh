mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s3_pytest6.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/s3_pytest6.txt
tail -25 gpurun_out/s3_pytest6.txt
timeout 900 python scripts/explore_spgemm.py --scale 20 --steps 2 --skip-check > gpurun_out/s3_explore20d.txt 2>&1
tail -12 gpurun_out/s3_explore20d.txt
