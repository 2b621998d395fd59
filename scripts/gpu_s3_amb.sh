python scripts/time_amb_convert.py 2>&1 | tail -5
timeout 600 python -m pytest tests/test_amb_gpu.py tests/test_drivers_gpu.py -x -q 2>&1 | tail -3
