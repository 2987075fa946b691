# 8-GPU validation + measurements (one gpurun --gpus 8 call)
set -x
export ODB_LOGLEVEL=WARNING
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29541 tests/dist_workers/outer_equiv.py > gpurun_out/equiv8.log 2>&1; grep -E "OK|FAIL" gpurun_out/equiv8.log | tail -9
timeout 400 $TR --master-port 29561 profiles/outer_sync_bench.py --iters 4 > gpurun_out/outer8.log 2>&1; grep "^{" gpurun_out/outer8.log
ODB_FUSED_OUTER_NO_MULTIMEM=1 timeout 200 $TR --master-port 29562 profiles/outer_sync_bench.py --iters 4 --models 150m --labels fused_fp32,fused_bf16 --no-ref > gpurun_out/outer8_p2p.log 2>&1; grep "^{" gpurun_out/outer8_p2p.log
timeout 300 $TR --master-port 29543 bench.py --gpus 8 --steps 4 --warmup 3 > gpurun_out/bench8.log 2>&1; tail -1 gpurun_out/bench8.log
