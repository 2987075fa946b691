#!/bin/bash
# round-2 GPU call G (8 GPUs): correctness of every outer transport at 8 ranks, outer-sync bench vs the reference statement
# sequence (BASELINE config #5), headline bench ours / reference at 8 GPUs, BASELINE config #4 (1B, 4 workers x 2 GPUs ZeRO-2)
# and config #3 (H = 50)
mkdir -p gpurun_out
export PYTHONPATH=$PWD
export ODB_LOGLEVEL=WARNING
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29541 tests/dist_workers/outer_equiv.py > gpurun_out/g_equiv8.log 2>&1; grep -E "OK|FAIL" gpurun_out/g_equiv8.log | tail -14
timeout 500 $TR --master-port 29561 profiles/outer_sync_bench.py --iters 5 --models 150m,1b --labels fused_fp32,fused_fp32_repl,fused_bf16,nccl_flat_fp32 > gpurun_out/g_outer8.jsonl 2> gpurun_out/g_outer8.err; cat gpurun_out/g_outer8.jsonl
for ctas in 120 250; do
ODB_OUTER_COMM_CTAS=$ctas timeout 200 $TR --master-port 29562 profiles/outer_sync_bench.py --iters 5 --models 150m --labels fused_fp32 --no-ref > gpurun_out/g_outer8_ctas$ctas.jsonl 2>/dev/null; echo "comm ctas $ctas"; cat gpurun_out/g_outer8_ctas$ctas.jsonl
done
timeout 400 $TR --master-port 29543 bench.py --gpus 8 --steps 4 --warmup 3 > gpurun_out/g_bench8_ours.json 2> gpurun_out/g_bench8_ours.err; tail -c 2500 gpurun_out/g_bench8_ours.json
timeout 500 $TR --master-port 29544 bench.py --impl reference --gpus 8 --steps 3 --warmup 3 > gpurun_out/g_bench8_ref.json 2> gpurun_out/g_bench8_ref.err; tail -c 700 gpurun_out/g_bench8_ref.json
# BASELINE config #3: 8 workers, H = 50 (two outer steps inside the timed window)
timeout 600 $TR --master-port 29545 bench.py --gpus 8 --steps 100 --warmup 5 --local-steps 50 --no-e2e > gpurun_out/g_bench8_h50.json 2> gpurun_out/g_bench8_h50.err; tail -c 1500 gpurun_out/g_bench8_h50.json
# BASELINE config #2 (H = 500) at a reduced batch (32 x 1024 tokens per worker-step, one micro-batch) so that 1000 steps and two
# real 8-GPU outer syncs fit in a minute; the full-batch figure follows from ms/step and the outer sync time above
timeout 400 $TR --master-port 29547 bench.py --gpus 8 --batch 32 --micro-batch 32 --steps 1000 --warmup 5 --local-steps 500 --no-e2e > gpurun_out/g_bench8_h500_b32.json 2> gpurun_out/g_bench8_h500_b32.err; tail -c 1200 gpurun_out/g_bench8_h500_b32.json
# BASELINE config #4: Llama-1B, 4 DiLoCo workers x 2 GPUs (ZeRO-2 inside a worker: fused ZeRO step + fused outer step)
timeout 500 $TR --master-port 29546 -m opendiloco_b200.train_fsdp --path-model 1b --fake-data --sharding-strategy _HYBRID_SHARD_ZERO2 \
  --per-device-train-batch-size 16 --total-batch-size 2048 --hv.local-steps 4 --hv.galaxy-size 4 --max-steps 9 \
  --metric-logger-type dummy --project gpurun_out/g_1b_zero2.pkl --no-torch-compile > gpurun_out/g_1b_zero2.log 2>&1; tail -3 gpurun_out/g_1b_zero2.log
python - <<'PY'
import pickle
try:
    d = pickle.load(open("gpurun_out/g_1b_zero2.pkl", "rb"))
    for m in d: print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in m.items() if k in ("step", "Loss", "time_taken", "tokens_per_second", "num_peers")})
except Exception as e:
    print("no 1b metrics:", e)
PY
