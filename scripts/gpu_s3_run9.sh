mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_spgemm_gpu.py tests/test_configs_gpu.py -x -q > gpurun_out/s3_pytest9.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/s3_pytest9.txt
tail -4 gpurun_out/s3_pytest9.txt
timeout 900 python scripts/explore_spgemm.py --scale 20 --steps 2 --skip-check --sweep "no_fork=0;no_fork=1" > gpurun_out/s3_explore20f.txt 2>&1
grep -E "options|step 1|bitmap" gpurun_out/s3_explore20f.txt
