"""BASELINE.json config #5: outer pseudo-gradient sync, parameter vector 150M -> 1B, fp32 / bf16, at N GPUs,
next to the reference's torch.distributed path.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29561 \
        profiles/outer_sync_bench.py [--models 150m,1b] [--iters 5]

"reference" = the exact statement sequence of train_diloco_torch.py:340-353 on tensors with the model's shapes: per
parameter {pageable H2D of the offloaded copy, subtract, NCCL all_reduce(AVG), re-point}, torch SGD(nesterov) step,
zero_grad, and the D2H clone of every parameter (get_offloaded_param).  Timed with CUDA events + barriers, max over ranks.
Effective GB/s = 4 * P bytes / time (the fp32 parameter vector once), the figure BASELINE.md uses.
"""
import argparse
import json
import os
import sys
from functools import partial

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opendiloco_b200.models.arena import llama_layout  # noqa: E402
from opendiloco_b200.models.config import LlamaConfig  # noqa: E402
from opendiloco_b200.optim.fused import FusedAdamW  # noqa: E402
from opendiloco_b200.parallel import comm  # noqa: E402
from opendiloco_b200.parallel.compression import get_compression  # noqa: E402
from opendiloco_b200.parallel.diloco import DiLoCoOptimizer  # noqa: E402
from opendiloco_b200.parallel.swarm import DHT  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--models", default="150m,1b")
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--labels", default="fused_fp32,fused_fp32_repl,fused_bf16,nccl_flat_fp32,blockwise8bit")
ap.add_argument("--no-ref", action="store_true")
a = ap.parse_args()

comm.init_distributed("nccl")
rank, world = dist.get_rank(), dist.get_world_size()
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))


def sync():
    torch.cuda.synchronize()
    dist.barrier(device_ids=[dev.index])
    torch.cuda.synchronize()


def timed(fn, iters):
    ts = []
    for _ in range(iters):
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        sync()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ts.append(float(t.item()))
    return min(ts), sorted(ts)[len(ts) // 2]


results = []
for name in a.models.split(","):
    cfg = LlamaConfig.from_pretrained(name)
    shapes = [shape for group in llama_layout(cfg) for _, shape in group]
    P = sum(int(torch.tensor(s).prod()) for s in shapes)
    torch.manual_seed(rank)

    # ------------------------------------------------------------------ reference statement sequence
    params = [torch.nn.Parameter(torch.randn(s, device=dev) * 0.02) for s in shapes]
    outer = torch.optim.SGD(params, lr=0.7, momentum=0.9, nesterov=True)
    offloaded = [p.data.detach().clone().to("cpu") for p in params]

    def reference_outer():
        nonlocal_off = offloaded
        for po, p in zip(nonlocal_off, params):
            pod = po.data.to(p.device)
            p.grad = pod - p.data
            dist.all_reduce(tensor=p.grad, op=dist.ReduceOp.AVG)
            p.data = pod
        outer.step()
        outer.zero_grad()
        nonlocal_off[:] = [p.data.detach().clone().to("cpu") for p in params]

    if a.no_ref:
        ref_min = float("nan")
    else:
        reference_outer()
        ref_min, ref_med = timed(reference_outer, max(2, a.iters // 2))
    del params, outer, offloaded
    torch.cuda.empty_cache()

    # ------------------------------------------------------------------ ours (flat arena of the same size)
    row = {"model": name, "params": P, "n_gpus": world, "reference_ms": ref_min, "reference_GBps": 4 * P / ref_min / 1e6}
    # fused_fp32 = sharded in-place kernel (ZeRO-1 outer optimizer, all-reduce on the master weights themselves);
    # fused_fp32_repl = the replicated-update pipelined kernel of round 1 (ODB_OUTER_SHARDED=0)
    for label, comp, fused in [("fused_fp32", None, True), ("fused_fp32_repl", None, True), ("fused_bf16", "bf16", True),
                               ("nccl_flat_fp32", None, False), ("blockwise8bit", "blockwise8bit", False)]:
        if label not in a.labels.split(","):
            continue
        os.environ["ODB_OUTER_SHARDED"] = "0" if label == "fused_fp32_repl" else "1"
        flat = torch.nn.Parameter(torch.randn(((P + 16383) // 16384) * 16384, device=dev) * 0.02)
        dist.broadcast(flat.data, src=0)          # workers of a real run start from identical weights (N1)
        opt = DiLoCoOptimizer(dht=DHT(start=True), batch_size=1, num_inner_steps=1, params=[flat],
                              outer_optimizer=partial(torch.optim.SGD, lr=0.7, momentum=0.9, nesterov=True),
                              inner_optimizer=partial(FusedAdamW, lr=0.0), grad_compression=get_compression(comp),
                              fused_collective=fused, timeout_waiting_for_peers=None)
        opt.timeout_waiting_for_peers = None      # pure kernel time: skip the store handshake
        used = opt._fused is not None

        def ours():
            flat.data.add_(1e-3)                  # a non-zero pseudo-gradient
            opt._update_global_epoch()

        ours()
        t_min, t_med = timed(ours, a.iters)
        # subtract the perturbation kernel (measured alone)
        p_min, _ = timed(lambda: flat.data.add_(1e-3), 3)
        t = max(t_min - p_min, 1e-3)
        row[label + "_ms"] = t
        row[label + "_GBps"] = 4 * P / t / 1e6
        row[label + "_fused_kernel"] = used
        if used:
            row[label + "_mode"] = "sharded" if opt._fused.sharded else ("pipelined" if opt._fused.pipelined else "sequential")
        if used and opt._fused.sharded:
            # the background part of the sharded form: owners re-replicate their momentum slab (NCCL all-gather, side stream)
            def regather():
                opt._fused._regather_pending = True
                opt._fused.start_momentum_regather()
                opt._fused.wait_momentum()

            regather()
            row[label + "_background_momentum_regather_ms"] = timed(regather, 3)[0]
        if used and opt._fused.phase_times_us() is not None:
            row[label + "_phases_us"] = opt._fused.phase_times_us()
        opt.shutdown()
        del opt, flat
        torch.cuda.empty_cache()
    results.append(row)
    if rank == 0:
        print(json.dumps(row), flush=True)

dist.barrier(device_ids=[dev.index])
comm.shutdown_distributed()
