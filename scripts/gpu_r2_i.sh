#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q -k "straggler_with_the_fused or fused_zero_step_kernel" > gpurun_out/i_tests.log 2>&1
echo "rc=$?"; tail -5 gpurun_out/i_tests.log; grep -E "FAIL" gpurun_out/i_tests.log | grep -v "^FAILED\|SOME" | head -10 | cut -c1-200
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 4 --steps 3 --warmup 3 --no-e2e > gpurun_out/i_bench4.json 2> gpurun_out/i_bench4.err
echo "bench rc=$?"; tail -c 1600 gpurun_out/i_bench4.json; tail -3 gpurun_out/i_bench4.err
