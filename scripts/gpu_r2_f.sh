#!/bin/bash
# round-2 GPU call F (4 GPUs): straggler policies on NCCL / on the fused NVLink kernels, elastic NO_WAIT rounds, 2 x 2 hybrid
mkdir -p gpurun_out
export PYTHONPATH=$PWD
nvidia-smi -L | wc -l
timeout 1700 python -m pytest tests/test_multi_gpu.py -m gpu -q -k "straggler or elastic or hybrid or outer_step_transports or fused_zero" > gpurun_out/f_tests_multi4.log 2>&1
echo "multi-gpu (4) tests rc=$?"; tail -40 gpurun_out/f_tests_multi4.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29561 profiles/outer_sync_bench.py --models 150m --iters 5 --no-ref --labels fused_fp32,fused_fp32_repl,fused_bf16,nccl_flat_fp32 > gpurun_out/f_outer_sync_4gpu.jsonl 2> gpurun_out/f_outer_sync_4gpu.err
cat gpurun_out/f_outer_sync_4gpu.jsonl
