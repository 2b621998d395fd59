set -x
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'num_bitmap|sym_bitmap' -s 2 -c 2 \
   -o gpurun_out/prof_bitmap_v2_s18 python scripts/explore_spgemm.py --scale 18 --steps 2 --skip-check > gpurun_out/ncu_full18_v2.log 2>&1
tail -3 gpurun_out/ncu_full18_v2.log
