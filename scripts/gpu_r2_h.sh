#!/bin/bash
# round-2 GPU call H (2 GPUs): fused ZeRO step equivalence + training, sharded outer kernel phase profile and comm-CTA sweep
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q -x -k "fused_zero" > gpurun_out/h_tests_zero.log 2>&1
echo "zero tests rc=$?"; tail -25 gpurun_out/h_tests_zero.log | cut -c1-250
for ctas in 16 32 48 64 100; do
ODB_OUTER_STAMPS=1 ODB_OUTER_COMM_CTAS=$ctas timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29563 profiles/outer_sync_bench.py --models 150m --iters 5 --no-ref --labels fused_fp32 > gpurun_out/h_outer_ctas_$ctas.jsonl 2>/dev/null
echo "comm ctas $ctas:"; python -c "
import json;d=json.loads(open('gpurun_out/h_outer_ctas_$ctas.jsonl').read().strip().splitlines()[-1]);print(d['fused_fp32_ms'], d.get('fused_fp32_phases_us'))"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 profiles/outer_sync_bench.py --models 150m --iters 5 --no-ref > gpurun_out/h_outer_sync_2gpu.jsonl 2> gpurun_out/h_outer_sync_2gpu.err
cat gpurun_out/h_outer_sync_2gpu.jsonl
