mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_configs_gpu.py -x -q -k "row_ranges or rerun" 2>&1 | tail -3
timeout 600 python scripts/explore_spgemm.py --scale 20 --dtype f64 --steps 2 --skip-check > gpurun_out/s3_explore20_f64.txt 2>&1
grep -E "step|bitmap|cta" gpurun_out/s3_explore20_f64.txt
