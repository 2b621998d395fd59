mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29513 scripts/probe_multicast.py > gpurun_out/probe_mc_$1.txt 2>&1; echo "probe rc=$?"; grep -v "^$" gpurun_out/probe_mc_$1.txt | grep -v "OMP_NUM\|\*\*\*" | tail -12
