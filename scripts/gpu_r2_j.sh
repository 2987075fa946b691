#!/bin/bash
# round-2 GPU call J (1 GPU): full GPU suite on the final tree, smoke(), RMSNorm-backward A/B, headline bench (+ reference arms)
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/j_tests_gpu.log 2>&1
echo "gpu suite rc=$?"; tail -4 gpurun_out/j_tests_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python - <<'PY'
import torch, os
from opendiloco_b200.ops import kernels as K
T,h=32768,1024
BF=torch.bfloat16
dy=torch.randn(T,h,device='cuda').to(BF); x=torch.randn(T,h,device='cuda').to(BF); w=torch.ones(h,device='cuda',dtype=BF)
rstd=torch.rand(T,device='cuda')+0.5; dres=torch.randn(T,h,device='cuda').to(BF); out=torch.empty_like(dres); dw=torch.zeros(h,device='cuda')
flush=torch.empty(256<<20,dtype=torch.uint8,device='cuda')
def t(fn,n=7):
    for _ in range(3): fn()
    ts=[]
    for _ in range(n):
        flush.zero_(); a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b)*1e3)
    return sorted(ts)[len(ts)//2]
us=t(lambda: K.rmsnorm_bwd(dy,x,w,rstd,dres,out,dw))
print(f"rmsnorm_bwd column-owner kernel: {us:.1f} us  ({(4*T*h*2)/us/1e3:.0f} GB/s of 4 x T x h x 2 B)")
PY
ODB_RMSNORM_BWD_V1=1 python - <<'PY'
import torch
from opendiloco_b200.ops import kernels as K
T,h=32768,1024
BF=torch.bfloat16
dy=torch.randn(T,h,device='cuda').to(BF); x=torch.randn(T,h,device='cuda').to(BF); w=torch.ones(h,device='cuda',dtype=BF)
rstd=torch.rand(T,device='cuda')+0.5; dres=torch.randn(T,h,device='cuda').to(BF); out=torch.empty_like(dres); dw=torch.zeros(h,device='cuda')
flush=torch.empty(256<<20,dtype=torch.uint8,device='cuda')
def t(fn,n=7):
    for _ in range(3): fn()
    ts=[]
    for _ in range(n):
        flush.zero_(); a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b)*1e3)
    return sorted(ts)[len(ts)//2]
us=t(lambda: K.rmsnorm_bwd(dy,x,w,rstd,dres,out,dw))
print(f"rmsnorm_bwd warp-per-row kernel (round 1): {us:.1f} us")
PY
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/j_bench.json 2> gpurun_out/j_bench.err; tail -c 1500 gpurun_out/j_bench.json
ODB_RMSNORM_BWD_V1=1 timeout 600 python bench.py --steps 8 --warmup 3 --no-e2e > gpurun_out/j_bench_rmsv1.json 2> gpurun_out/j_bench_rmsv1.err
python -c "
import json
for f in ('j_bench','j_bench_rmsv1'):
    d=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1]); print(f, round(d['value']), d['ms_per_step'], d['clocks']['sm_mhz'], d['gpu_launches'])"
timeout 600 python bench.py --impl reference --steps 4 --warmup 3 > gpurun_out/j_bench_ref.json 2> gpurun_out/j_bench_ref.err; tail -c 400 gpurun_out/j_bench_ref.json
