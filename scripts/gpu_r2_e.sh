#!/bin/bash
# round-2 GPU call E (2 GPUs): fused ZeRO step test, outer kernels retest, outer-sync bench after tuning
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 1200 python -m pytest tests/test_multi_gpu.py -m gpu -q -x -k "fused_zero or outer_step_transports" > gpurun_out/e_tests_multi.log 2>&1
echo "tests rc=$?"; tail -30 gpurun_out/e_tests_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 profiles/outer_sync_bench.py --models 150m,1b --iters 5 --no-ref > gpurun_out/e_outer_sync_2gpu.jsonl 2> gpurun_out/e_outer_sync_2gpu.err
echo "outer bench rc=$?"; cat gpurun_out/e_outer_sync_2gpu.jsonl; grep -v WARNING gpurun_out/e_outer_sync_2gpu.err | tail -5
for ctas in 100 148 250; do
ODB_OUTER_COMM_CTAS=$ctas timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29563 profiles/outer_sync_bench.py --models 150m --iters 5 --no-ref --labels fused_fp32 > gpurun_out/e_outer_ctas_$ctas.jsonl 2>/dev/null
echo "comm ctas $ctas:"; python -c "
import json;d=json.loads(open('gpurun_out/e_outer_ctas_$ctas.jsonl').read().strip().splitlines()[-1]);print(d['fused_fp32_ms'])"
done
