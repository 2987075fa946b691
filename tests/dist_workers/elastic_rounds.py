"""Run under torchrun with 3 ranks (gloo on CPU, or NCCL on 3 GPUs with ODB_TEST_DEVICE=cuda).  AllReduceStrategy.NO_WAIT with a straggler: outer step 1 splits into
the round of the two punctual workers and a solo round of the late one; outer step 2 is a full round again and repairs
the drift with a state-averaging round.  Every transition is checked against values computed by hand."""
import os
import sys
import time
from functools import partial

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from opendiloco_b200.parallel import comm  # noqa: E402
from opendiloco_b200.parallel.diloco import AllReduceStrategy, DiLoCoOptimizer  # noqa: E402
from opendiloco_b200.parallel.swarm import DHT  # noqa: E402

cuda = os.environ.get("ODB_TEST_DEVICE", "cpu") == "cuda"
comm.init_distributed("nccl" if cuda else "gloo")
rank, world = dist.get_rank(), dist.get_world_size()
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0))) if cuda else torch.device("cpu")
assert world == 3
LATE, H = 2, 2
os.environ["ODB_FAULT_INJECT"] = f"{LATE}:1:5.0"          # worker 2 reaches outer step 1 five seconds late
p = torch.nn.Parameter(torch.zeros(64, device=dev))
opt = DiLoCoOptimizer(dht=DHT(start=True), run_id="elastic", batch_size=1, num_inner_steps=H, params=[p],
                      outer_optimizer=partial(torch.optim.SGD, lr=1.0, momentum=0.0),
                      inner_optimizer=partial(torch.optim.SGD, lr=1.0), all_reduce_strategy=AllReduceStrategy.NO_WAIT,
                      matchmaking_time=2.0, averaging_timeout=30.0, fused_collective=False)
ok = True


def check(name, cond):
    global ok
    ok = ok and bool(cond)
    print(f"[rank {rank}] {name}: {'OK' if cond else 'FAIL'}", flush=True)


# the subset butterfly itself: ranks 0 and 2 average a 1003-element vector (uneven slices), rank 1 is not involved
if rank in (0, 2):
    v = torch.arange(1003, dtype=torch.float32, device=dev) * (rank + 1)
    comm.p2p_all_reduce_mean_(v, [0, 2], dist.group.WORLD)
    check("p2p subset mean", torch.allclose(v, torch.arange(1003, dtype=torch.float32, device=dev) * 2.0))
    w = torch.full((5,), float(rank), dtype=torch.bfloat16, device=dev)   # fewer elements than members x 8: empty slices
    comm.p2p_all_reduce_mean_(w, [0, 2], dist.group.WORLD)
    check("p2p subset mean, tiny bf16", torch.allclose(w.float(), torch.full((5,), 1.0, device=dev)))


def inner_steps(scale):
    """H plain-SGD steps with gradient = scale  =>  theta_local moves by -H * scale."""
    for _ in range(H):
        p.grad = torch.full_like(p, scale)
        opt.step()
        opt.zero_grad()


# epoch 0: everybody punctual.  pseudo-gradient of worker r = H * (r + 1); outer SGD(lr 1) subtracts the mean
inner_steps(rank + 1.0)
check("epoch0 full round", opt.last_round_members == [0, 1, 2] and opt.local_epoch == 1)
theta0 = -H * (1 + 2 + 3) / 3
check("epoch0 value", torch.allclose(p.data, torch.full_like(p, theta0)))

# epoch 1: worker 2 is late -> round [0, 1], then worker 2 alone
inner_steps(rank + 1.0)
if rank == LATE:
    check("epoch1 solo round", opt.last_round_members == [LATE])
    expect = theta0 - H * 3.0
else:
    check("epoch1 partial round", opt.last_round_members == [0, 1])
    expect = theta0 - H * (1 + 2) / 2
check("epoch1 value", torch.allclose(p.data, torch.full_like(p, expect)) and opt.local_epoch == 2)

# epoch 2: the punctual workers idle long enough for the straggler to catch up, so everybody meets again (the late worker
# now arrives first and leads the round).  Each worker applies the mean pseudo-gradient to ITS theta_outer, then the drift
# repair averages theta_outer over all three workers.
if rank != LATE:
    time.sleep(4.0)      # punctual workers end epoch 1 at ~2 s, the straggler at ~5 s: both reach outer step 2 within 2 s
inner_steps(rank + 1.0)
check("epoch2 full round", opt.last_round_members == [0, 1, 2] and opt.local_epoch == 3)
pre = [theta0 - H * 1.5, theta0 - H * 1.5, theta0 - H * 3.0]
final = sum(x - H * 2.0 for x in pre) / 3
check("epoch2 value (drift repaired)", torch.allclose(p.data, torch.full_like(p, final), atol=1e-5))
gathered = [torch.zeros_like(p.data) for _ in range(world)]
dist.all_gather(gathered, p.data)
check("all workers agree", all(torch.equal(gathered[0], g) for g in gathered))

# epochs 3-7: worker 2 stalls for four outer steps of the others (their published records then show epoch 6: a record is
# refreshed at inner steps).  Back again it first closes its own round of epoch 3 (the swarm is ahead), is then still more
# than one epoch behind, asks the board for the swarm state and receives it (point to point) from the leader of the
# others' next round
if rank == LATE:
    time.sleep(10.0)
    p.grad = torch.full_like(p, 1.0)
    opt.step()                                  # sees the swarm ahead: closes its own (solo) round of epoch 3
    opt.zero_grad()
    check("laggard caught one epoch on its own", opt.local_epoch == 4)
    p.grad = torch.full_like(p, 1.0)
    opt.step()                                  # still > 1 epoch behind: state request, served after the others' next round
    opt.zero_grad()
    check("laggard adopted the swarm epoch", opt.local_epoch == 8)
else:
    for _ in range(4):
        inner_steps(rank + 1.0)                 # rounds [0, 1]: the leader waits matchmaking_time for worker 2 each time
    check("punctual workers moved on", opt.local_epoch == 7 and opt.last_round_members == [0, 1])
    time.sleep(5.0)
    inner_steps(rank + 1.0)                     # the leader of this round serves the pending state request
    check("epoch after serving", opt.local_epoch == 8)
dist.all_gather(gathered, p.data)
check("laggard holds the swarm parameters", all(torch.equal(gathered[0], g) for g in gathered))

# epoch 8: one more FULL round after the resync.  The laggard adopted the state of ONE peer, so it arrives flagged as
# drifted; the round leader turns that into ``repair`` for every member - all three run the state-averaging collective
# (a decision taken from process-local flags would leave the punctual workers out of it and dead-lock / corrupt the round)
if rank != LATE:
    time.sleep(1.0)
inner_steps(rank + 1.0)
check("epoch8 full round", opt.last_round_members == [0, 1, 2] and opt.local_epoch == 9)
dist.all_gather(gathered, p.data)
check("all workers agree after the post-resync round", all(torch.equal(gathered[0], g) for g in gathered))
check("drift flag cleared", not opt._drifted)
opt.shutdown()
dist.barrier()
print(f"[rank {rank}] {'ALL OK' if ok else 'SOME FAILED'}", flush=True)
comm.shutdown_distributed()
sys.exit(0 if ok else 1)
