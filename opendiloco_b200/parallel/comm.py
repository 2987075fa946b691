"""Process topology and collectives.

One process per GPU, ``torch.distributed`` (NCCL over NVLink 5 / NVSwitch; gloo on CPU) for the plumbing.  The job is
a grid ``galaxy_size x gpus_per_worker``:

    rank = world_rank * gpus_per_worker + local_rank
    inner group  = the ranks of one DiLoCo worker   (gradient reduce / ZeRO sharding every step)
    outer group  = same local_rank across workers   (pseudo-gradient all-reduce every H steps)

which is the reference's "galaxy" of workers (train_fsdp.py:91-92,151-156) collapsed onto one NCCL world instead of
one torchrun + hivemind swarm per worker (SURVEY.md §1 process topology, §5.8).  Workers may also be launched as
separate ``torchrun`` jobs that rendezvous on one TCP store (``--hv.initial-peers tcp://host:port``), mirroring
run_training.sh.

Every collective here is issued ONCE on a flat buffer (the reference issues one per parameter tensor:
train_diloco_torch.py:255,345; train_fsdp.py:410-413).
"""
from __future__ import annotations

import datetime
import os
from dataclasses import dataclass

import torch
import torch.distributed as dist

TIMEOUT_NCCL_MINUTES = int(os.environ.get("TIMEOUT_NCCL_MINUTES", 120))  # reference: train_fsdp.py:64


def default_backend() -> str:
    return "nccl" if torch.cuda.is_available() else "gloo"


def init_distributed(backend: str | None = None, init_method: str | None = None, rank: int | None = None,
                     world_size: int | None = None) -> None:
    """Idempotent process-group bring-up from the torchrun environment (reference: ddp_setup, train_fsdp.py:70-72)."""
    if dist.is_initialized():
        return
    backend = backend or default_backend()
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    kwargs: dict = dict(backend=backend, timeout=datetime.timedelta(minutes=TIMEOUT_NCCL_MINUTES))
    if backend == "nccl":
        torch.cuda.set_device(local_rank)
        kwargs["device_id"] = torch.device("cuda", local_rank)
    if init_method is not None:
        kwargs.update(init_method=init_method, rank=rank, world_size=world_size)
    elif "RANK" not in os.environ:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29512")
        kwargs.update(rank=0, world_size=1)
    dist.init_process_group(**kwargs)


def shutdown_distributed() -> None:
    if dist.is_initialized():
        dist.destroy_process_group()


@dataclass
class Topology:
    rank: int
    world_size: int
    local_rank: int                 # rank inside the DiLoCo worker
    gpus_per_worker: int
    world_rank: int                 # index of this DiLoCo worker
    galaxy_size: int                # number of DiLoCo workers
    inner_group: dist.ProcessGroup | None
    outer_group: dist.ProcessGroup | None

    @property
    def is_worker_leader(self) -> bool:
        return self.local_rank == 0


def build_topology(galaxy_size: int | None = None, gpus_per_worker: int | None = None) -> Topology:
    """Carve the default world into (workers x gpus-per-worker).  Groups are created collectively by every rank."""
    if not dist.is_initialized():
        return Topology(0, 1, 0, 1, 0, 1, None, None)
    world, rank = dist.get_world_size(), dist.get_rank()
    if galaxy_size is None and gpus_per_worker is None:
        galaxy_size, gpus_per_worker = world, 1
    elif galaxy_size is None:
        galaxy_size = world // gpus_per_worker
    elif gpus_per_worker is None:
        gpus_per_worker = world // galaxy_size
    assert galaxy_size * gpus_per_worker == world, (galaxy_size, gpus_per_worker, world)
    wr, lr = divmod(rank, gpus_per_worker)
    inner = outer = None
    if gpus_per_worker > 1:
        for w in range(galaxy_size):
            g = dist.new_group(list(range(w * gpus_per_worker, (w + 1) * gpus_per_worker)))
            if w == wr:
                inner = g
    if galaxy_size > 1:
        if gpus_per_worker == 1:
            outer = dist.group.WORLD
        else:
            for l in range(gpus_per_worker):
                g = dist.new_group(list(range(l, world, gpus_per_worker)))
                if l == lr:
                    outer = g
    return Topology(rank, world, lr, gpus_per_worker, wr, galaxy_size, inner, outer)


def group_size(group) -> int:
    if group is None or not dist.is_initialized():
        return 1
    return dist.get_world_size(group)


def all_reduce_avg_(buf: torch.Tensor, group) -> torch.Tensor:
    """In-place mean over the group on ONE flat buffer.  NCCL has a native AVG; gloo gets SUM + scale."""
    n = group_size(group)
    if n == 1:
        return buf
    if buf.is_cuda:
        dist.all_reduce(buf, op=dist.ReduceOp.AVG, group=group)
    else:
        work = buf if buf.dtype in (torch.float32, torch.float64) else buf.float()
        dist.all_reduce(work, op=dist.ReduceOp.SUM, group=group)
        work.div_(n)
        if work is not buf:
            buf.copy_(work)
    return buf


def all_reduce_sum_(buf: torch.Tensor, group) -> torch.Tensor:
    if group_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return buf


def broadcast_(buf: torch.Tensor, src_rank_in_group: int, group) -> torch.Tensor:
    """Flat broadcast (reference N1/N6: one broadcast per tensor, train_diloco_torch.py:253-255, train_fsdp.py:410-413)."""
    if group_size(group) > 1:
        src = dist.get_global_rank(group, src_rank_in_group) if group is not dist.group.WORLD else src_rank_in_group
        dist.broadcast(buf, src=src, group=group)
    return buf


def barrier(group=None) -> None:
    if dist.is_initialized() and dist.get_world_size() > 1:
        if torch.cuda.is_available() and dist.get_backend() == "nccl":
            dist.barrier(group=group, device_ids=[torch.cuda.current_device()])
        else:
            dist.barrier(group=group)
