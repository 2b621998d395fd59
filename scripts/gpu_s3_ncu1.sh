mkdir -p gpurun_out
timeout 1700 ncu --set full --clock-control none --import-source on -k regex:"sym_bitmap|num_bitmap" -c 2 -o gpurun_out/s3_prof_bitmap_s20 -f python scripts/explore_spgemm.py --scale 20 --steps 1 --skip-check > gpurun_out/s3_ncu1.log 2>&1
tail -5 gpurun_out/s3_ncu1.log
ls -la gpurun_out/*.ncu-rep
