#!/bin/bash
# round-2 GPU call A: new kernels' tests, full GPU suite, LCE micro-bench, bench A/B (fused LCE on/off)
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests/test_lce_gpu.py tests/test_tc_gemm_gpu.py -x -q > gpurun_out/a_tests_new.log 2>&1
echo "new tests rc=$?" | tee -a gpurun_out/a_tests_new.log
tail -5 gpurun_out/a_tests_new.log
timeout 600 python profiles/lce_bench.py > gpurun_out/a_lce_bench.log 2>&1; tail -15 gpurun_out/a_lce_bench.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/a_tests_gpu.log 2>&1
echo "gpu suite rc=$?" | tee -a gpurun_out/a_tests_gpu.log
tail -5 gpurun_out/a_tests_gpu.log
ODB_LCE_FUSED=0 timeout 600 python bench.py --steps 6 --warmup 3 --no-e2e > gpurun_out/a_bench_lce0.json 2> gpurun_out/a_bench_lce0.err
timeout 600 python bench.py --steps 6 --warmup 3 --no-e2e > gpurun_out/a_bench_lce1.json 2> gpurun_out/a_bench_lce1.err
ODB_LCE_FUSED=0 timeout 600 python bench.py --steps 6 --warmup 3 --no-e2e > gpurun_out/a_bench_lce0b.json 2>> gpurun_out/a_bench_lce0.err
timeout 600 python bench.py --steps 6 --warmup 3 --no-e2e > gpurun_out/a_bench_lce1b.json 2>> gpurun_out/a_bench_lce1.err
for f in gpurun_out/a_bench_lce*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d.get('clocks'))"; done
