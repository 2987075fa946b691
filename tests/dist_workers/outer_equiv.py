"""Run under torchrun (nccl on GPUs, gloo on CPU).  Checks that N DiLoCo workers with different data reproduce the
reference's outer-step semantics (train_diloco_torch.py:336-353: AVG of theta_outer-theta_local, Nesterov SGD, reset)
for every transport: flat collective, fused symmetric-memory kernel (fp32 / bf16 window), compressed butterfly."""
import os
import sys
from functools import partial

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from opendiloco_b200.optim.fused import FusedAdamW  # noqa: E402
from opendiloco_b200.parallel import comm  # noqa: E402
from opendiloco_b200.parallel.compression import get_compression  # noqa: E402
from opendiloco_b200.parallel.diloco import DiLoCoOptimizer  # noqa: E402
from opendiloco_b200.parallel.swarm import DHT  # noqa: E402

comm.init_distributed()
rank, world = dist.get_rank(), dist.get_world_size()
cuda = torch.cuda.is_available()
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0))) if cuda else torch.device("cpu")
H, STEPS = 3, 9


def make_model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(64, 256), torch.nn.Tanh(), torch.nn.Linear(256, 128)).to(dev)


def data_for(r):
    g = torch.Generator().manual_seed(100 + r)
    return [(torch.randn(16, 64, generator=g).to(dev), torch.randn(16, 128, generator=g).to(dev)) for _ in range(STEPS)]


def oracle():
    """All workers simulated in one process with plain torch optimizers."""
    models = [make_model() for _ in range(world)]
    inner = [torch.optim.AdamW(m.parameters(), lr=1e-2, weight_decay=0.1, betas=(0.9, 0.95)) for m in models]
    outer = [torch.optim.SGD(m.parameters(), lr=0.7, momentum=0.9, nesterov=True) for m in models]
    off = [[p.data.clone() for p in m.parameters()] for m in models]
    datas = [data_for(r) for r in range(world)]
    for s in range(1, STEPS + 1):
        for w, m in enumerate(models):
            x, y = datas[w][s - 1]
            torch.nn.functional.mse_loss(m(x), y).backward()
            torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
            inner[w].step()
            inner[w].zero_grad()
        if s % H == 0:
            deltas = [[po - p.data for po, p in zip(off[w], models[w].parameters())] for w in range(world)]
            mean = [sum(d[i] for d in deltas) / world for i in range(len(deltas[0]))]
            for w, m in enumerate(models):
                for po, p, g in zip(off[w], m.parameters(), mean):
                    p.grad = g.clone()
                    p.data = po
                outer[w].step()
                outer[w].zero_grad()
                off[w] = [p.data.clone() for p in m.parameters()]
    return torch.cat([p.data.reshape(-1) for p in models[rank].parameters()])


def ours(compression=None, fused=None, env=None):
    for k, v in (env or {}).items():
        os.environ[k] = v
    m = make_model()
    dht = DHT(start=True)
    opt = DiLoCoOptimizer(dht=dht, batch_size=16, num_inner_steps=H, params=m.parameters(),
                          outer_optimizer=partial(torch.optim.SGD, lr=0.7, momentum=0.9, nesterov=True),
                          inner_optimizer=partial(FusedAdamW, lr=1e-2, weight_decay=0.1, betas=(0.9, 0.95), max_grad_norm=1.0),
                          grad_compression=get_compression(compression), fused_collective=fused, timeout_waiting_for_peers=60.0,
                          matchmaking_time=1.0)
    for s in range(STEPS):
        x, y = data_for(rank)[s]
        torch.nn.functional.mse_loss(m(x), y).backward()
        opt.step()
        opt.zero_grad()
    assert opt.local_epoch == STEPS // H
    used_fused = opt._fused is not None
    mode = ("sharded" if opt._fused.sharded else "pipelined" if opt._fused.pipelined else "sequential") if used_fused else "-"
    out = torch.cat([p.data.reshape(-1) for p in m.parameters()])
    # the FULL outer momentum must be identical on every worker (the sharded kernel re-replicates it in the background)
    mom = opt.state_dict()["state_dict_outer"]["state"]
    mom = torch.cat([mom[k]["momentum_buffer"].reshape(-1).to(dev) for k in sorted(k for k in mom if isinstance(mom[k], dict))])
    # shadow (bf16 compute weights) must track the master weights after the outer step
    fv = opt.inner_optimizer.fv
    shadow_ok = fv.shadow is None or torch.equal(fv.shadow[:fv.flat.numel()].float(), fv.flat.to(fv.shadow.dtype).float())
    opt.shutdown()
    for k in (env or {}):
        os.environ.pop(k, None)
    return out, used_fused, mode, mom, shadow_ok


ref = oracle()
results = {}
cases = [("flat", None, False, None)]
if cuda:
    cases += [("fused_fp32", None, True, None),                                   # sharded in-place kernel when NVLS is there
              ("fused_fp32_repl", None, True, {"ODB_OUTER_SHARDED": "0"}),        # replicated-update pipelined kernel
              ("fused_fp32_seq", None, True, {"ODB_OUTER_SHARDED": "0", "ODB_OUTER_PIPELINED": "0"}),
              ("fused_fp32_p2p", None, True, {"ODB_FUSED_OUTER_NO_MULTIMEM": "1"}),    # peer loads / stores, no multicast
              ("fused_bf16", "bf16", True, None)]
cases += [(c, c, False, None) for c in ("fp16", "bf16", "scaled-fp16", "uniform8bit", "quantile8bit", "blockwise8bit")]
tol = {"flat": 2e-6, "fused_fp32": 2e-6, "fused_fp32_repl": 2e-6, "fused_fp32_seq": 2e-6, "fused_fp32_p2p": 2e-6,
       "fused_bf16": 2e-3, "fp16": 1e-3, "bf16": 3e-3, "scaled-fp16": 1e-3,
       "uniform8bit": 8e-2, "quantile8bit": 8e-2, "blockwise8bit": 2e-2}
ok = True
for name, comp, fused, env in cases:
    out, used, mode, mom, shadow_ok = ours(comp, fused, env)
    err = (out - ref).abs().max().item()
    # all workers must hold identical parameters (and outer momentum) right after an outer step
    gathered = [torch.empty_like(out) for _ in range(world)]
    dist.all_gather(gathered, out)
    spread = max((g - gathered[0]).abs().max().item() for g in gathered)
    gm = [torch.empty_like(mom) for _ in range(world)]
    dist.all_gather(gm, mom)
    mspread = max((g - gm[0]).abs().max().item() for g in gm)
    good = err < tol[name] and spread < 1e-6 and mspread < 1e-6 and shadow_ok and \
        (not fused or used or os.environ.get("ODB_ALLOW_NO_FUSED"))
    ok &= good
    if rank == 0:
        print(f"{name:16s} err={err:.3e} spread={spread:.1e} momentum_spread={mspread:.1e} shadow={'ok' if shadow_ok else 'BAD'} "
              f"fused={mode} {'OK' if good else 'FAIL'}", flush=True)
dist.barrier()
comm.shutdown_distributed()
sys.exit(0 if ok else 1)
