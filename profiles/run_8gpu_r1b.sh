# 8-GPU validation + measurements, second pass of round 1 (one `gpurun --gpus 8` call)
set -x
export ODB_LOGLEVEL=WARNING
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29541 tests/dist_workers/outer_equiv.py > gpurun_out/equiv8.log 2>&1; grep -E "OK|FAIL" gpurun_out/equiv8.log | tail -9
# outer step: pipelined fused kernel (default) vs phase-sequential vs flat NCCL, 150M; then the 1B vector
timeout 300 $TR --master-port 29561 profiles/outer_sync_bench.py --iters 4 --models 150m,1b --labels fused_fp32,fused_bf16,nccl_flat_fp32 --no-ref > gpurun_out/outer8.log 2>&1; grep "^{" gpurun_out/outer8.log
ODB_OUTER_PIPELINED=0 timeout 200 $TR --master-port 29562 profiles/outer_sync_bench.py --iters 4 --models 150m --labels fused_fp32,fused_bf16 --no-ref > gpurun_out/outer8_seq.log 2>&1; grep "^{" gpurun_out/outer8_seq.log
timeout 300 $TR --master-port 29543 bench.py --gpus 8 --steps 4 --warmup 3 > gpurun_out/bench8.log 2>&1; tail -1 gpurun_out/bench8.log
# BASELINE config #4: Llama-1B, 4 DiLoCo workers x 2 GPUs (ZeRO-2 inside a worker); tokens/s per worker from the metric log
timeout 400 $TR --master-port 29544 -m opendiloco_b200.train_fsdp --path-model 1b --fake-data --sharding-strategy _HYBRID_SHARD_ZERO2 \
  --per-device-train-batch-size 16 --total-batch-size 2048 --hv.local-steps 4 --hv.galaxy-size 4 --max-steps 9 \
  --metric-logger-type dummy --project gpurun_out/1b_zero2.pkl --no-torch-compile > gpurun_out/1b_zero2.log 2>&1; tail -3 gpurun_out/1b_zero2.log
python - <<'PY'
import pickle
try:
    d = pickle.load(open("gpurun_out/1b_zero2.pkl", "rb"))
    for m in d: print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in m.items() if k in ("step", "Loss", "time_taken", "tokens_per_second", "num_peers")})
except Exception as e:
    print("no 1b metrics:", e)
PY
