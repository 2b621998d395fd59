mkdir -p gpurun_out
timeout 1500 python scripts/explore_spgemm.py --scale 20 --steps 2 --skip-check --phases --sweep "sym_bitmap_min=512,num_bitmap_min=256" > gpurun_out/s3_phases.txt 2>&1
cat gpurun_out/s3_phases.txt
