#!/bin/bash
# round-2 GPU call D (1 GPU): tightened model tests, idle-gap, CUDA-graph A/B, ncu of the bandwidth kernels, reference-fsdp arm,
# loss parity (reference vs ours, 1000 steps)
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_lce_gpu.py -m gpu -q -s > gpurun_out/d_tests_model.log 2>&1
echo "model tests rc=$?"; grep -E "worst|passed|failed|Error" gpurun_out/d_tests_model.log | tail -12
timeout 600 python profiles/idle_gap.py > gpurun_out/d_idle_gap.log 2>&1; tail -1 gpurun_out/d_idle_gap.log
timeout 600 python profiles/idle_gap.py --graph > gpurun_out/d_idle_gap_graph.log 2>&1; tail -1 gpurun_out/d_idle_gap_graph.log
timeout 600 python bench.py --steps 6 --warmup 3 --no-e2e > gpurun_out/d_bench_nograph.json 2> gpurun_out/d_bench_nograph.err
ODB_CUDA_GRAPH=1 timeout 600 python bench.py --steps 6 --warmup 3 --no-e2e > gpurun_out/d_bench_graph.json 2> gpurun_out/d_bench_graph.err
timeout 600 python bench.py --steps 6 --warmup 3 --no-e2e > gpurun_out/d_bench_nograph2.json 2>> gpurun_out/d_bench_nograph.err
ODB_CUDA_GRAPH=1 timeout 600 python bench.py --steps 6 --warmup 3 > gpurun_out/d_bench_graph2.json 2>> gpurun_out/d_bench_graph.err
for f in gpurun_out/d_bench_*.json; do echo $f; python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d.get('clocks',{}).get('sm_mhz'), (d.get('e2e') or {}).get('value'))"; done
tail -3 gpurun_out/d_bench_graph.err
NCU="ncu --set full --clock-control none --import-source on"
timeout 900 $NCU -k regex:"grad_sqnorm|adamw_step|nesterov_outer|rmsnorm_|embedding_|cast_" -s 8 -c 8 -f -o gpurun_out/d_ncu_bw python profiles/ncu_one.py bandwidth > gpurun_out/d_ncu_bw.log 2>&1
echo "ncu bw rc=$?"
timeout 1200 python bench.py --impl reference-fsdp --steps 6 --warmup 3 > gpurun_out/d_bench_ref_fsdp.json 2> gpurun_out/d_bench_ref_fsdp.err
echo "ref-fsdp rc=$?"; tail -c 800 gpurun_out/d_bench_ref_fsdp.json; tail -5 gpurun_out/d_bench_ref_fsdp.err
timeout 900 python profiles/loss_parity.py --impl ours --steps 1000 --out gpurun_out/d_parity_ours.jsonl > gpurun_out/d_parity_ours.log 2>&1
echo "parity ours rc=$?"; tail -2 gpurun_out/d_parity_ours.log
timeout 1500 python profiles/loss_parity.py --impl reference --steps 1000 --out gpurun_out/d_parity_ref.jsonl > gpurun_out/d_parity_ref.log 2>&1
echo "parity ref rc=$?"; tail -2 gpurun_out/d_parity_ref.log
python profiles/loss_parity.py --compare gpurun_out/d_parity_ref.jsonl gpurun_out/d_parity_ours.jsonl | tee gpurun_out/d_parity_compare.txt
