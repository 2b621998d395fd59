set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_spgemm_gpu.py -x -q > gpurun_out/pytest_spgemm.log 2>&1; tail -5 gpurun_out/pytest_spgemm.log
timeout 300 python scripts/explore_spgemm.py --scale 18 --steps 2 > gpurun_out/explore18.log 2>&1; tail -12 gpurun_out/explore18.log
timeout 900 python scripts/explore_spgemm.py --scale 20 --steps 2 --skip-check > gpurun_out/explore20.log 2>&1; tail -12 gpurun_out/explore20.log
