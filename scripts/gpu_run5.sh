set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_spgemm_gpu.py -x -q > gpurun_out/pytest_spgemm.log 2>&1; tail -3 gpurun_out/pytest_spgemm.log
timeout 900 python scripts/explore_spgemm.py --scale 20 --steps 2 --skip-check > gpurun_out/explore20_p256.log 2>&1; tail -9 gpurun_out/explore20_p256.log
NSP_LIB_PATH=$PWD/nsparse_b200/lib/libnsparse_b200_p128.so timeout 900 python scripts/explore_spgemm.py --scale 20 --steps 2 --skip-check > gpurun_out/explore20_p128.log 2>&1; tail -9 gpurun_out/explore20_p128.log
timeout 300 python scripts/explore_spgemm.py --scale 18 --steps 2 --skip-check > gpurun_out/explore18.log 2>&1; tail -9 gpurun_out/explore18.log
