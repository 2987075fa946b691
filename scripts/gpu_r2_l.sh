#!/bin/bash
# round-2 GPU call L (1 GPU): final tree - full GPU suite, smoke, bench
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/l_tests_gpu.log 2>&1
echo "gpu suite rc=$?"; tail -4 gpurun_out/l_tests_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/l_bench.json 2> gpurun_out/l_bench.err; tail -c 900 gpurun_out/l_bench.json; tail -2 gpurun_out/l_bench.err
