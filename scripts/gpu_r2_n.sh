#!/bin/bash
# round-2 GPU call N (1 GPU): programmatic dependent launch - correctness (GPU suite) and A/B
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/n_tests_gpu.log 2>&1
echo "gpu suite rc=$?"; tail -4 gpurun_out/n_tests_gpu.log
ODB_PDL=0 timeout 400 python bench.py --steps 6 --warmup 3 --no-e2e > gpurun_out/n_bench_pdl0.json 2> gpurun_out/n_bench_pdl0.err
timeout 400 python bench.py --steps 6 --warmup 3 --no-e2e > gpurun_out/n_bench_pdl1.json 2> gpurun_out/n_bench_pdl1.err
ODB_PDL=0 timeout 400 python bench.py --steps 6 --warmup 3 --no-e2e > gpurun_out/n_bench_pdl0b.json 2>> gpurun_out/n_bench_pdl0.err
timeout 400 python bench.py --steps 6 --warmup 3 --no-e2e > gpurun_out/n_bench_pdl1b.json 2>> gpurun_out/n_bench_pdl1.err
python -c "
import json
for f in ('n_bench_pdl0','n_bench_pdl1','n_bench_pdl0b','n_bench_pdl1b'):
    try:
        d=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1]); print(f, round(d['value']), round(d['ms_per_step'],2), d['clocks']['sm_mhz'])
    except Exception as e: print(f, 'failed', e)"
tail -3 gpurun_out/n_bench_pdl1.err
