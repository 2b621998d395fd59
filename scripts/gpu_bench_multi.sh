N=$1; shift
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 --no-e2e "$@" > gpurun_out/bench_g${N}_x.json 2> gpurun_out/bench_g$N.err; echo "bench$N $@ rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/bench_g${N}_x.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['config'].get('per_rank'))"
tail -2 gpurun_out/bench_g$N.err | cut -c1-200
