#!/bin/bash
# gpurun -- bash scripts/gpu_round_end.sh : what the driver runs at round end, plus the profiler evidence of the round
R=r2
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${R}_pytest_final.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/${R}_pytest_final.txt
tail -4 gpurun_out/${R}_pytest_final.txt
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${R}_bench_ref.json 2> gpurun_out/${R}_bench_final.err; echo "ref rc=$?"
timeout 900 python bench.py > gpurun_out/${R}_bench_final.json 2>> gpurun_out/${R}_bench_final.err; echo "bench rc=$?"
tail -3 gpurun_out/${R}_bench_final.err | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-spmv --no-ref-gpu > gpurun_out/${R}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"sym_bitmap|num_bitmap|num_hash_kernel" -c 6 -o gpurun_out/${R}_prof_spgemm -f python scripts/explore_spgemm.py --scale 20 --steps 1 --skip-check > gpurun_out/${R}_ncu_spgemm.log 2>&1; echo "ncu spgemm rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"amb_spmv" -c 2 -o gpurun_out/${R}_prof_spmv -f python scripts/spmv_c3.py 4096 > gpurun_out/${R}_ncu_spmv.log 2>&1; echo "ncu spmv rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_smoke.txt 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/${R}_smoke.txt
# racecheck of the cursor-search staging after the __syncwarp fix (profiles/r2_sanitizer/)
python scripts/make_sanitize_inputs.py /tmp > /dev/null 2>&1
NSP_OPTIONS=sym_bitmap_min=64,num_bitmap_min=64,sym_window_shift=16,num_window_shift=16,num_cap=128,no_seg=1 timeout 600 compute-sanitizer --tool racecheck --print-limit 20 oracle/_ref/dump_spgemm_ours_s /tmp/sanitize_a_s.bin /tmp/sanitize_b_s.bin /tmp/san_recheck.bin 0 > gpurun_out/${R}_racecheck_forced_search_after_fix.log 2>&1; grep -E "RACECHECK SUMMARY|hazard" gpurun_out/${R}_racecheck_forced_search_after_fix.log | head -5
