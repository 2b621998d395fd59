mkdir -p gpurun_out
timeout 1700 ncu --set full --clock-control none --import-source on -k regex:"sym_bitmap|num_bitmap" -c 2 -o gpurun_out/s3_prof_bitmap_s20_v4 -f python scripts/explore_spgemm.py --scale 20 --steps 1 --skip-check --opt sym_bitmap_min=512 --opt num_bitmap_min=256 > gpurun_out/s3_ncu3.log 2>&1
tail -3 gpurun_out/s3_ncu3.log
