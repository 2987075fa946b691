#!/bin/bash
# round-2 GPU call C (2 GPUs): multi-process tests over NCCL + outer-sync bench of every transport at 2 GPUs
mkdir -p gpurun_out
export PYTHONPATH=$PWD
nvidia-smi -L
timeout 1500 python -m pytest tests/test_multi_gpu.py -m gpu -q -x > gpurun_out/c_tests_multi.log 2>&1
echo "multi-gpu tests rc=$?"; tail -30 gpurun_out/c_tests_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 profiles/outer_sync_bench.py --models 150m,1b --iters 5 --no-ref > gpurun_out/c_outer_sync_2gpu.jsonl 2> gpurun_out/c_outer_sync_2gpu.err
echo "outer bench rc=$?"; cat gpurun_out/c_outer_sync_2gpu.jsonl; tail -5 gpurun_out/c_outer_sync_2gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 2 --steps 4 --warmup 3 --no-e2e > gpurun_out/c_bench_2gpu.json 2> gpurun_out/c_bench_2gpu.err
echo "bench rc=$?"; tail -c 1200 gpurun_out/c_bench_2gpu.json; tail -3 gpurun_out/c_bench_2gpu.err
