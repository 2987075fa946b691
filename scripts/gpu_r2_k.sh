#!/bin/bash
# round-2 GPU call K (8 GPUs): final-tree confirmation - outer sync with the new owner-CTA default (96) and 64, headline bench
mkdir -p gpurun_out
export PYTHONPATH=$PWD
export ODB_LOGLEVEL=WARNING
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29561 profiles/outer_sync_bench.py --iters 6 --models 150m,1b --labels fused_fp32 --no-ref > gpurun_out/k_outer8_default.jsonl 2>/dev/null; cat gpurun_out/k_outer8_default.jsonl
ODB_OUTER_COMM_CTAS=64 timeout 200 $TR --master-port 29562 profiles/outer_sync_bench.py --iters 6 --models 150m --labels fused_fp32 --no-ref > gpurun_out/k_outer8_ctas64.jsonl 2>/dev/null; cat gpurun_out/k_outer8_ctas64.jsonl
timeout 400 $TR --master-port 29543 bench.py --gpus 8 --steps 4 --warmup 3 > gpurun_out/k_bench8_ours.json 2> gpurun_out/k_bench8_ours.err; tail -c 1500 gpurun_out/k_bench8_ours.json
