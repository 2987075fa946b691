#!/bin/bash
# Local multi-worker DiLoCo launcher (positional contract of the reference's open_diloco/run_training.sh:28-36):
#
#   ./run_training.sh <N workers> <GPUs per worker> <initial_peer|auto> [train_fsdp flags ...]
#   ./run_training.sh 4 2 auto --per-device-train-batch-size 16 --total-batch-size 512 --hv.local-steps 10 --fake-data
#
# Each worker is its own torchrun over a disjoint CUDA_VISIBLE_DEVICES slice and logs to logs/log<i>.  Workers meet on
# ONE rendezvous store ("auto" => tcp://127.0.0.1:$DILOCO_PORT hosted by worker 0) and form a single NCCL world of
# N x GPUs ranks, so the outer all-reduce runs over NVLink - there is no DHT daemon to start.
set -euo pipefail

gpu_slice() {  # <gpus per worker> <worker index> -> "a,b,c"
    local n=$1 i=$2 s=$(( $1 * $2 ))
    seq -s ',' "$s" $(( s + n - 1 ))
}

if [ "$#" -lt 3 ]; then
    echo "Usage: $0 <N> <num_gpu> <initial_peer|auto> [additional train_fsdp args]"
    exit 1
fi
N=$1; NUM_GPU=$2; INITIAL_PEER=$3; shift 3
PORT=${DILOCO_PORT:-29400}
if [ "$INITIAL_PEER" = "auto" ]; then INITIAL_PEER="tcp://127.0.0.1:${PORT}"; fi
echo "Initial peer: $INITIAL_PEER"
mkdir -p logs
HERE="$(cd "$(dirname "$0")" && pwd)"

# optional native membership board (heartbeats / progress records visible to scripts/swarm_status.py): DILOCO_BOARD_PORT=29500
if [ -n "${DILOCO_BOARD_PORT:-}" ]; then
    PYTHONPATH="$HERE" python -m opendiloco_b200.parallel.rendezvous --serve "$DILOCO_BOARD_PORT" > logs/board.log 2>&1 &
    export ODB_BOARD="odb://127.0.0.1:${DILOCO_BOARD_PORT}"
fi

for i in $(seq 0 $(( N - 1 ))); do
    extra_env=""
    if [ "$i" -gt 0 ]; then export_wandb="WANDB_MODE=disabled"; else export_wandb=""; fi
    env $export_wandb CUDA_VISIBLE_DEVICES="$(gpu_slice "$NUM_GPU" "$i")" PYTHONPATH="$HERE" \
        torchrun --nproc_per_node="$NUM_GPU" --nnodes=1 --master-addr 127.0.0.1 --master-port $(( PORT + 1 + i )) \
        -m opendiloco_b200.train_fsdp --hv.initial-peers "$INITIAL_PEER" "$@" --hv.world-rank "$i" --hv.galaxy-size "$N" \
        > "logs/log$i" 2>&1 &
    if [ "$i" -eq 0 ]; then sleep 2; fi
done
# DILOCO_WAIT=1: block until every worker has exited and return the first failure (CI / tests); default: follow worker 0's
# log like the reference script does.
if [ -n "${DILOCO_WAIT:-}" ]; then
    rc=0
    for pid in $(jobs -p); do wait "$pid" || rc=$?; done
    exit $rc
fi
tail -f logs/log0
