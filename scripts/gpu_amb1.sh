set -x
mkdir -p gpurun_out/golden
python tests/golden/make_amb_golden.py gpurun_out/golden > gpurun_out/make_amb_golden.log 2>&1; tail -8 gpurun_out/make_amb_golden.log
cp gpurun_out/golden/*.npz tests/golden/ 2>/dev/null
timeout 900 python -m pytest tests/test_amb_oracle.py -x -q -k fixtures > gpurun_out/pytest_amb_fixtures.log 2>&1; tail -15 gpurun_out/pytest_amb_fixtures.log
timeout 1200 python -m pytest tests/test_amb_gpu.py -x -q > gpurun_out/pytest_amb_gpu.log 2>&1; tail -30 gpurun_out/pytest_amb_gpu.log
python - <<'PY' > gpurun_out/spmv_c3.log 2>&1
import time, numpy as np, torch, sys
sys.path.insert(0,'.')
import nsparse_b200 as ns
from nsparse_b200 import gen
ctx=ns.Context(0)
for n in (1024, 4096):
    lap=gen.laplacian5_csr(n); lap.memcpy()
    x=torch.rand(lap.N,dtype=torch.float64,device='cuda')
    t=time.time(); amb=ns.csr2amb(lap,ctx=ctx); torch.cuda.synchronize(); tc=time.time()-t
    y=torch.empty(lap.M,dtype=torch.float64,device='cuda')
    for _ in range(5): ns.spmv_amb(amb,x,out=y,ctx=ctx)
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): ns.spmv_amb(amb,x,out=y,ctx=ctx)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/50
    alg=lap.nnz*12+4*(lap.M+1)+8*(lap.N+lap.M)
    # check vs torch csr
    A=torch.sparse_csr_tensor(lap.d_rpt.long(),lap.d_col.long(),lap.d_val,size=(lap.M,lap.N))
    y0=A@x
    err=((y-y0).abs().max()/y0.abs().max()).item()
    print(f"n={n} conv={tc:.3f}s seg={amb.seg_size} bs={amb.block_size} c_size={amb.c_size} nnz_amb={amb.nnz} ms={ms:.4f} GFLOPS={2*lap.nnz/ms/1e6:.1f} GB/s_alg={alg/ms/1e6:.1f} err={err:.2e}")
PY
cat gpurun_out/spmv_c3.log
