mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_final.txt
tail -4 gpurun_out/pytest_final.txt
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2>> gpurun_out/bench_final.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"sym_bitmap|num_bitmap" -c 3 -o gpurun_out/prof_final -f python scripts/explore_spgemm.py --scale 20 --steps 1 --skip-check > gpurun_out/ncu_final.log 2>&1; echo "ncu full rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.txt
