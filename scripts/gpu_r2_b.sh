#!/bin/bash
# round-2 GPU call B (1 GPU): GPU suite, LCE micro-bench, ncu --set full of the LCE + bandwidth kernels, launch list, bench
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/b_tests_gpu.log 2>&1
echo "gpu suite rc=$?"; tail -8 gpurun_out/b_tests_gpu.log
timeout 600 python profiles/lce_bench.py > gpurun_out/b_lce_bench.log 2>&1; tail -12 gpurun_out/b_lce_bench.log
NCU="ncu --set full --clock-control none --import-source on"
timeout 900 $NCU -k regex:"gemm2_bf16_tn_kernel|lce_|wgrad_kernel" -s 5 -c 5 -f -o gpurun_out/b_ncu_lce python profiles/ncu_one.py lce > gpurun_out/b_ncu_lce.log 2>&1
echo "ncu lce rc=$?"
timeout 900 $NCU -s 8 -c 8 -f -o gpurun_out/b_ncu_bw python profiles/ncu_one.py bandwidth > gpurun_out/b_ncu_bw.log 2>&1
echo "ncu bw rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 700 --csv --log-file gpurun_out/b_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e > gpurun_out/b_launches_bench.log 2>&1
echo "launch list rc=$?"
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err
tail -c 1500 gpurun_out/b_bench.json
ls -la gpurun_out | head -40
