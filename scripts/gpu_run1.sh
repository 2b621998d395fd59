set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt
nproc >> gpurun_out/gpu.txt; free -g >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python scripts/explore_spgemm.py --scale 18 --steps 2 > gpurun_out/explore18.log 2>&1; tail -20 gpurun_out/explore18.log
timeout 900 python scripts/explore_spgemm.py --scale 20 --steps 2 > gpurun_out/explore20.log 2>&1; tail -20 gpurun_out/explore20.log
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench20.log 2>&1; tail -2 gpurun_out/bench20.log
