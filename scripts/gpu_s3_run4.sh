mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_spgemm_gpu.py -x -q > gpurun_out/s3_pytest4.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/s3_pytest4.txt
tail -15 gpurun_out/s3_pytest4.txt
timeout 900 python scripts/explore_spgemm.py --scale 18 --steps 2 > gpurun_out/s3_explore18c.txt 2>&1
tail -14 gpurun_out/s3_explore18c.txt
timeout 1500 python scripts/explore_spgemm.py --scale 20 --steps 2 --skip-check --sweep "num_window_shift=19;sym_bitmap_min=512,num_bitmap_min=256;sym_bitmap_min=4096,num_bitmap_min=2048;sym_bitmap_min=16384,num_bitmap_min=8192" > gpurun_out/s3_explore20c.txt 2>&1
cat gpurun_out/s3_explore20c.txt
