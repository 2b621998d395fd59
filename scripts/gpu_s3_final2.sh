mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s3_pytest_final2.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/s3_pytest_final2.txt
tail -4 gpurun_out/s3_pytest_final2.txt
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/s3_bench_final2.json 2> gpurun_out/s3_bench_final2.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/s3_bench_final2.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['spmv']['conversion_s'], d['spmv']['conversion_first_call_s'], d['spmv']['value'])"
