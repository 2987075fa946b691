"""Run under torchrun with >= 2 ranks; gloo on CPU, or NCCL on GPUs with ODB_TEST_DEVICE=cuda.

AllReduceStrategy.WAIT_FOR_ALL with a straggler (ODB_FAULT_INJECT): the punctual workers wait ``timeout_waiting_for_peers``
for it, then "skip the slowest peer" and average among themselves (reference: hivemind_diloco.py:584-607); the late worker
closes the epoch on its own when it finally arrives; the next full round repairs the drift with a state-averaging round
and every worker ends up with identical parameters.  Values are checked against numbers computed by hand."""
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
LATE, H = world - 1, 2
os.environ["ODB_FAULT_INJECT"] = f"{LATE}:1:7.0"          # the last worker reaches outer step 1 seven seconds late
p = torch.nn.Parameter(torch.zeros(4096, device=dev))
opt = DiLoCoOptimizer(dht=DHT(start=True), run_id="straggler", batch_size=1, num_inner_steps=H, params=[p],
                      outer_optimizer=partial(torch.optim.SGD, lr=1.0, momentum=0.5, nesterov=True),
                      inner_optimizer=partial(torch.optim.SGD, lr=1.0), all_reduce_strategy=AllReduceStrategy.WAIT_FOR_ALL,
                      timeout_waiting_for_peers=2.5, matchmaking_time=1.0, averaging_timeout=30.0,
                      fused_collective=os.environ.get("ODB_TEST_FUSED", "0") == "1")
ok = True


def check(name, cond):
    global ok
    ok = ok and bool(cond)
    print(f"[rank {rank}] {name}: {'OK' if cond else 'FAIL'}", flush=True)


def inner_steps(scale):
    for _ in range(H):
        p.grad = torch.full_like(p, scale)
        opt.step()
        opt.zero_grad()


def nesterov(theta, buf, d, lr=1.0, mu=0.5):
    buf = mu * buf + d
    return theta - lr * (d + mu * buf), buf


everyone = list(range(world))
punctual = everyone[:-1]
g = [float(r + 1) for r in everyone]                       # worker r's gradient scale: pseudo-gradient = H * (r + 1)

# epoch 0: full round
inner_steps(g[rank])
check("epoch0 full round", opt.last_round_members == everyone and opt.local_epoch == 1)
th0, b0 = nesterov(0.0, 0.0, H * sum(g) / world)
check("epoch0 value", torch.allclose(p.data, torch.full_like(p, th0)))

# epoch 1: the straggler misses the window -> the others skip it; it closes the epoch alone afterwards
t0 = time.perf_counter()
inner_steps(g[rank])
if rank == LATE:
    check("epoch1 solo round of the straggler", opt.last_round_members == [LATE])
    th1, b1 = nesterov(th0, b0, H * g[LATE])
else:
    waited = time.perf_counter() - t0
    check("epoch1 round without the straggler", opt.last_round_members == punctual)
    check("waited ~timeout_waiting_for_peers, not for the straggler", 2.0 < waited < 6.0)
    th1, b1 = nesterov(th0, b0, H * sum(g[:-1]) / len(punctual))
check("epoch1 value", torch.allclose(p.data, torch.full_like(p, th1), atol=1e-5) and opt.local_epoch == 2)

# epoch 2: the punctual workers idle until the straggler has caught up -> full round; theta_outer and the momentum are
# averaged over everybody afterwards (drift repair), so all workers agree again
if rank != LATE:
    time.sleep(6.0)
inner_steps(g[rank])
check("epoch2 full round", opt.last_round_members == everyone and opt.local_epoch == 3)
thp, bp = nesterov(th0, b0, H * sum(g[:-1]) / len(punctual))
thl, bl = nesterov(th0, b0, H * g[LATE])
d2 = H * sum(g) / world
outs = [nesterov(thp, bp, d2) for _ in punctual] + [nesterov(thl, bl, d2)]
final = sum(o[0] for o in outs) / world
check("epoch2 value (drift repaired)", torch.allclose(p.data, torch.full_like(p, final), atol=1e-4))
gathered = [torch.zeros_like(p.data) for _ in range(world)]
dist.all_gather(gathered, p.data)
check("all workers agree", all(torch.equal(gathered[0], x) for x in gathered))

# epoch 3: an ordinary full round, no repair needed any more
inner_steps(g[rank])
check("epoch3 full round", opt.last_round_members == everyone and opt.local_epoch == 4 and not opt._drifted)
dist.all_gather(gathered, p.data)
check("all workers still agree", all(torch.equal(gathered[0], x) for x in gathered))
opt.shutdown()
dist.barrier()
print(f"[rank {rank}] {'ALL OK' if ok else 'SOME FAILED'}", flush=True)
comm.shutdown_distributed()
sys.exit(0 if ok else 1)
