set -x
mkdir -p gpurun_out
# launch list of the bench command (single pass, no replay)
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench20.csv \
   python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-spmv > gpurun_out/ncu_bench20.log 2>&1
tail -2 gpurun_out/ncu_bench20.log | cut -c1-300
# full capture of the two dominant kernels at scale 18 (10 GB of C: replay stays cheap)
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'num_bitmap|sym_bitmap' -s 2 -c 2 \
   -o gpurun_out/prof_bitmap_s18 python scripts/explore_spgemm.py --scale 18 --steps 2 > gpurun_out/ncu_full18.log 2>&1
tail -5 gpurun_out/ncu_full18.log
ls -la gpurun_out
