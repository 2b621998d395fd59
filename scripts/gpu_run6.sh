mkdir -p gpurun_out
for d in 0 1 2 3 4 7; do
echo "debug=$d"
timeout 900 python scripts/explore_spgemm.py --scale 20 --steps 2 --skip-check --opt debug=$d 2>&1 | grep -E "num_bitmap"
done
