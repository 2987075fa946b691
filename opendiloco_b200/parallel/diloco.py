"""DiLoCo: inner optimizer every step, outer (pseudo-gradient) step every ``num_inner_steps``.

Public surface mirrors the reference's ``open_diloco.hivemind_diloco``:
``DiLoCoOptimizer``, ``DiLoCoGradAverager``, ``DiLoCoStateAverager``, ``DiloCoProgressTracker``, ``AllReduceStrategy``.
The machinery underneath is different by design:

  * theta_outer, the pseudo-gradient and the outer momentum are flat fp32 buffers RESIDENT IN HBM (the reference
    keeps them on the CPU in shared memory: hivemind_diloco.py:400, train_diloco_torch.py:132-135);
  * the pseudo-gradient all-reduce is ONE collective on the flat buffer over NCCL/NVLink, or - when an NVLink
    symmetric-memory window is available - a single fused kernel that computes theta_outer - theta_local, reduces it
    through the switch (multimem) and applies Nesterov in the same launch (``parallel/fused_outer.py``);
  * the outer SGD-Nesterov update, the theta_local reset and the bf16 shadow refresh are one kernel
    (``csrc/optim.cu``) instead of foreach SGD + 111 copies (train_diloco_torch.py:346-353).

Algorithm (reference: hivemind_diloco.py:483-558,570-679 ; train_diloco_torch.py:336-353):
    every step:      inner.step() ; samples += batch_size
    every H steps:   delta = theta_outer - theta_local ; delta <- mean over workers ;
                     theta_outer <- SGD_nesterov(theta_outer, delta) ; theta_local <- theta_outer ; epoch += 1
Inner AdamW moments and the LR schedule are NOT reset at outer steps.
"""
from __future__ import annotations

import contextlib
import os
import time
from enum import Enum
from typing import Callable, Iterable

import torch
import torch.distributed as dist

from ..ops import kernels as K
from ..optim.fused import FlatView, flatten_params
from ..utils.logger import get_logger
from . import comm
from .swarm import (DHT, GlobalTrainingProgress, LocalTrainingProgress, PerformanceEMA, StepControl, get_dht_time)

logger = get_logger()


class AllReduceStrategy(Enum):
    """When to trigger the pseudo-gradient averaging round (reference: hivemind_diloco.py:285-297).

    WAIT_FOR_ALL: wait (up to ``timeout_waiting_for_peers``) until every worker finished its local steps.
    NO_WAIT:      the fastest worker triggers the round after ``matchmaking_time``; workers that have not arrived by
                  then are left out of this round and re-synchronise from their peers afterwards.
    """

    WAIT_FOR_ALL = "WAIT_FOR_ALL"
    NO_WAIT = "NO_WAIT"


DEFAULT_TIMEOUT_WAITING_FOR_PEERS = 600


# ====================================================================================================== progress tracker
class DiloCoProgressTracker:
    """Local/global progress bookkeeping (reference: hivemind_diloco.py:174-282 on top of hivemind.ProgressTracker).

    An epoch (= one outer step) is ready when THIS worker accumulated ``target_batch_size = batch_size * H`` samples
    or when the swarm is already ahead.  Peers publish (epoch, samples, samples/s) through the c10d store so that
    ``global_progress`` / ETA mean the same thing as in the reference.
    """

    def __init__(self, batch_size: int, num_inner_steps: int, *, dht: DHT | None = None, prefix: str = "diloco",
                 target_batch_size: int | None = None, min_refresh_period: float = 0.5, max_refresh_period: float = 2.0,
                 default_refresh_period: float = 1.0, performance_ema_alpha: float = 0.1, publish: bool = False,
                 peer_ttl: float | None = None, **_ignored):
        self.batch_size, self.num_inner_steps = batch_size, num_inner_steps
        self.dht, self.prefix = dht, prefix
        self.target_batch_size = target_batch_size if target_batch_size is not None else batch_size * num_inner_steps
        self.min_refresh_period, self.max_refresh_period = min_refresh_period, max_refresh_period
        self.default_refresh_period = default_refresh_period
        self.performance_ema = PerformanceEMA(alpha=performance_ema_alpha)
        self.local_progress = LocalTrainingProgress(peer_id=dht.peer_id if dht else "worker-0")
        self._global_epoch_hint = 0
        self._paused = False
        self._publish = publish
        # a peer whose last record is older than this many seconds no longer counts as alive (the reference's DHT records
        # carry an expiration time for the same purpose: hivemind ProgressTracker, metadata_expiration); a worker that
        # died therefore shows up as num_peers < galaxy_size -> "Lost a diloco worker" (train_fsdp.py:440-446)
        self.peer_ttl = float(os.environ.get("ODB_PEER_TTL", 120.0)) if peer_ttl is None else float(peer_ttl)
        self.global_progress = self._make_global()

    # -- reference properties ------------------------------------------------------------------
    @property
    def num_peers(self) -> int:
        return self.dht.num_peers if self.dht is not None else 1

    @property
    def global_epoch(self) -> int:
        return max(self._global_epoch_hint, self.local_progress.epoch)

    @property
    def ready_to_update_epoch(self) -> bool:
        return (self.global_epoch > self.local_progress.epoch
                or self.local_progress.samples_accumulated >= self.target_batch_size)

    @property
    def estimated_next_update_time(self) -> float:
        """Seconds until this peer reaches its local steps (0 when ready).  The reference returns an absolute time in
        the ready case and a duration otherwise; callers compare it with a duration, so a
        duration is returned in both cases."""
        if self.ready_to_update_epoch:
            return 0.0
        remaining = max(0, self.target_batch_size - self.local_progress.samples_accumulated)
        return remaining / max(self.performance_ema.samples_per_second, 1e-9)

    @property
    def local_step(self) -> int:
        return self.local_progress.samples_accumulated // self.batch_size

    @property
    def real_step(self) -> int:
        return self.local_step + self.local_progress.epoch * self.num_inner_steps

    # -- updates -----------------------------------------------------------------------------
    def _make_global(self) -> GlobalTrainingProgress:
        now = get_dht_time()
        eta = self.estimated_next_update_time
        return GlobalTrainingProgress(epoch=self.global_epoch, samples_accumulated=self.local_progress.samples_accumulated,
                                      target_batch_size=self.target_batch_size, num_peers=self.num_peers, num_clients=0,
                                      eta_next_epoch=now + eta,
                                      next_fetch_time=now + min(max(eta, self.min_refresh_period), self.max_refresh_period))

    def report_local_progress(self, local_epoch: int, samples_accumulated: int, update_global_samples: bool = True) -> None:
        extra = samples_accumulated - self.local_progress.samples_accumulated
        if extra > 0 and not self._paused:
            self.performance_ema.update(task_size=extra)
        self.local_progress.epoch = local_epoch
        self.local_progress.samples_accumulated = samples_accumulated
        self.local_progress.samples_per_second = self.performance_ema.samples_per_second
        self.local_progress.time = get_dht_time()
        fetch_due = self._publish and get_dht_time() >= self.global_progress.next_fetch_time
        self.global_progress = self._make_global()
        if self._publish:
            self._publish_progress()
            if fetch_due:                     # hivemind refreshes the swarm view in a background thread at this cadence
                self.fetch_global_progress()

    def update_epoch(self, new_epoch: int | None = None) -> int:
        if new_epoch is None:
            new_epoch = self.local_progress.epoch + 1
        self._global_epoch_hint = max(self._global_epoch_hint, new_epoch)
        self.local_progress.epoch = new_epoch
        self.local_progress.samples_accumulated = 0
        self.performance_ema.reset_timer()
        self.global_progress = self._make_global()
        return new_epoch

    @contextlib.contextmanager
    def pause_updates(self):
        """Freeze throughput accounting (used around the outer step and checkpoint writes: hivemind_diloco.py:610,
        train_fsdp.py:478)."""
        self._paused = True
        try:
            with self.performance_ema.pause():
                yield
        finally:
            self._paused = False

    # -- optional cross-worker visibility through the rendezvous store ------------------------------
    def _publish_progress(self) -> None:
        store = self.dht.store() if self.dht is not None else None
        if store is None:
            return
        lp = self.local_progress
        store.set(f"{self.prefix}_progress/{lp.peer_id}",
                  f"{lp.epoch},{lp.samples_accumulated},{lp.samples_per_second:.6f},{lp.time:.3f}")

    def fetch_global_progress(self) -> GlobalTrainingProgress:
        """Read every peer's published record; ETA = max over peers of their remaining time (hivemind_diloco.py:248-256)."""
        store = self.dht.store() if self.dht is not None else None
        if store is None or not self._publish:
            return self.global_progress
        now = get_dht_time()
        eta, epoch, alive = 0.0, self.local_progress.epoch, 0
        for pid in self.dht.peer_ids():
            try:
                if not store.check([f"{self.prefix}_progress/{pid}"]):
                    continue
                e, s, sps, t_rec = store.get(f"{self.prefix}_progress/{pid}").decode().split(",")
            except Exception:
                continue
            if pid != self.local_progress.peer_id and now - float(t_rec) > self.peer_ttl:
                continue                      # stale record: the peer stopped reporting
            alive += 1
            epoch = max(epoch, int(e))
            if int(e) >= epoch:
                eta = max(eta, max(0, self.target_batch_size - int(s)) / max(float(sps), 1e-9))
        self._global_epoch_hint = max(self._global_epoch_hint, epoch)
        # the swarm status line hivemind prints per fetch (hivemind_diloco.py:269-272)
        logger.debug(f"{self.prefix} has taken {self.local_step} local steps. Peers: {alive}, epoch: {epoch}, "
                     f"steps: {self.real_step}. ETA: {eta:.2f}")
        self.global_progress = GlobalTrainingProgress(epoch, 0, self.target_batch_size, num_peers=max(alive, 1),
                                                      num_clients=0, eta_next_epoch=now + eta,
                                                      next_fetch_time=now + min(max(eta, self.min_refresh_period),
                                                                                self.max_refresh_period))
        return self.global_progress


# ====================================================================================================== gradient averager
class DiLoCoGradAverager:
    """Averages pseudo-gradients across workers; the averaged tensors ARE the ``.grad`` buffers of the outer
    ("offloaded") optimizer (reference: hivemind_diloco.py:61-171).

    ``flat=(theta_outer, delta, theta_local)`` is supplied by DiLoCoOptimizer when everything lives in flat buffers;
    stand-alone use with arbitrary parameters (reference test tests/test_diloco_hivemind.py:53-96) stages through a
    temporary flat buffer.
    """

    def __init__(self, main_parameters, offloaded_optimizer: torch.optim.Optimizer, *, dht: DHT | None = None,
                 prefix: str = "diloco_grad_averager", warn: bool = True, compression=None, flat=None,
                 min_matchmaking_time: float = 5.0, **kwargs):
        if kwargs.pop("client_mode", None):
            raise KeyError("client_mode is not supported in DiLoCoGradAverager")
        if "averaged_grads" in kwargs:
            raise KeyError("DiLoCoGradAverager does not support averaged_grads: it uses the offloaded optimizer gradients")
        if not isinstance(main_parameters, (list, tuple)):
            raise ValueError("main_parameters must be a list or tuple of parameters, not an iterator")
        self.main_parameters = list(main_parameters)
        self.offloaded_optimizer = offloaded_optimizer
        self.dht, self.prefix, self.warn = dht, prefix, warn
        self.compression = compression
        self.matchmaking_kwargs = {"min_matchmaking_time": min_matchmaking_time}
        self.local_samples_accumulated = 0
        self.local_times_accumulated = 0
        self._new_averaged_grads = False
        self._flat = flat
        self._averaged_grads = tuple(self._grads_from_optimizer())
        self.last_allreduce_seconds = 0.0

    @property
    def peer_id(self) -> str:
        return self.dht.peer_id if self.dht is not None else "worker-0"

    @property
    def group(self):
        return self.dht.group if self.dht is not None else None

    def _offloaded_params(self) -> list[torch.Tensor]:
        return [p for g in self.offloaded_optimizer.param_groups for p in g["params"]]

    def _grads_from_optimizer(self):
        for p in self._offloaded_params():
            if p.grad is None:
                p.grad = torch.zeros_like(p)
            yield p.grad

    @contextlib.contextmanager
    def get_tensors(self):
        yield self._averaged_grads

    # -- scheduling (matchmaking is a no-op on a static NVLink group; the handle keeps the reference's call pattern)
    def schedule_step(self, scheduled_time: float | None = None, **kwargs) -> StepControl:
        assert kwargs.get("weight") is None, "setting weight in schedule_step is not supported"
        return StepControl(scheduled_time=scheduled_time)

    @torch.no_grad()
    def compute_and_load_pseudo_grad_into_averager(self) -> None:
        """delta = theta_outer - theta_local into the outer optimizer's grad buffers (hivemind_diloco.py:158-167)."""
        if self._flat is not None:
            theta_outer, delta, theta_local = self._flat
            if delta.device == theta_local.device:
                K.pseudo_grad(theta_outer, theta_local, delta)
            else:
                delta.copy_(theta_outer - theta_local.to(theta_outer.device))
            return
        for opt_p, g, main_p in zip(self._offloaded_params(), self._averaged_grads, self.main_parameters):
            g.copy_(opt_p.data - main_p.detach().to(opt_p.device), non_blocking=True)

    @torch.no_grad()
    def _all_reduce(self, members: list[int] | None = None) -> None:
        group = self.group
        if comm.group_size(group) <= 1:
            return
        t0 = time.perf_counter()
        if members is not None and len(members) < comm.group_size(group):
            # elastic round (NO_WAIT): only the peers that formed this round exchange data, point to point
            if self._flat is not None:
                comm.p2p_all_reduce_mean_(self._flat[1], members, group)
            else:
                grads = list(self._averaged_grads)
                dev = comm_device(group, grads[0].device)
                flat = torch.cat([g.reshape(-1).to(dev, torch.float32) for g in grads])
                comm.p2p_all_reduce_mean_(flat, members, group)
                off = 0
                for g in grads:
                    g.copy_(flat[off:off + g.numel()].view_as(g))
                    off += g.numel()
            self.last_allreduce_seconds = time.perf_counter() - t0
            return
        if self._flat is not None:
            buf = self._flat[1]
            if self.compression is not None and not self.compression.is_identity:
                self.compression.all_reduce_mean_(buf, group)
            else:
                comm.all_reduce_avg_(_collective_view(buf), group)
                if not buf.is_cuda and dist.get_backend(group) == "nccl":   # CPU-offloaded theta_outer, NCCL group
                    raise RuntimeError("offload_device='cpu' needs a gloo group")
        else:
            grads = list(self._averaged_grads)
            dev = comm_device(group, grads[0].device)
            flat = torch.cat([g.reshape(-1).to(dev, torch.float32) for g in grads])
            if self.compression is not None and not self.compression.is_identity:
                self.compression.all_reduce_mean_(flat, group)
            else:
                comm.all_reduce_avg_(flat, group)
            off = 0
            for g in grads:
                g.copy_(flat[off:off + g.numel()].view_as(g))
                off += g.numel()
        self.last_allreduce_seconds = time.perf_counter() - t0

    def step(self, control: StepControl | None = None, timeout: float | None = None, wait: bool = True,
             members: list[int] | None = None, **kwargs):
        """Compute the pseudo-gradient, average it with the peers, leave the mean in the outer grads
        (hivemind_diloco.py:134-156).  Returns {peer_id: None} for every member of the round (or the control if
        ``wait=False``; the collective itself is stream-ordered, nothing runs on a background thread).
        ``members``: ranks (in the outer group) that formed this round; None = everybody."""
        if control is None:
            control = self.schedule_step(timeout=timeout, **kwargs)
        self.compute_and_load_pseudo_grad_into_averager()
        control.allow_allreduce()
        try:
            self._all_reduce(members)
            self._new_averaged_grads = True
            if members is not None:
                ids = [f"worker-{r}" for r in members]
            else:
                ids = self.dht.peer_ids() if self.dht is not None else [self.peer_id]
            control.set_result({pid: None for pid in ids})
        except BaseException as e:  # surfaced through control.result(), like an MPFuture
            control.set_exception(e)
        return control.result(timeout) if wait else control

    def notify_used_averaged_gradients(self) -> None:
        self._new_averaged_grads = False

    def shutdown(self) -> None:
        pass


def _collective_view(buf: torch.Tensor) -> torch.Tensor:
    return buf


def comm_device(group, fallback: torch.device) -> torch.device:
    if group is not None and dist.is_initialized() and dist.get_backend(group) == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return fallback


# ====================================================================================================== state averager
class DiLoCoStateAverager:
    """Owns theta_outer ("offloaded" parameters), the outer optimizer built on it, the inner optimizer and its LR
    scheduler (reference: hivemind_diloco.py:35-58 on top of hivemind.TrainingStateAverager).
    The scheduler is attached to the INNER optimizer only; the outer optimizer never sees it."""

    def __init__(self, *, params, optimizer: Callable, inner_optimizer: torch.optim.Optimizer, num_inner_steps: int,
                 scheduler: Callable | None = None, dht: DHT | None = None, prefix: str = "diloco_state_averager",
                 offload_device: str | torch.device | None = None, flat_view: FlatView | None = None,
                 average_state_every: int = 1, **_ignored):
        self.dht, self.prefix = dht, prefix
        self.inner_optimizer, self.num_inner_steps = inner_optimizer, num_inner_steps
        self.main_parameters = list(params)
        self.fv = flat_view if flat_view is not None else flatten_params(self.main_parameters)
        dev = self.fv.flat.device if offload_device is None else torch.device(offload_device)
        self.theta_local = self.fv.own(self.fv.flat)          # the slice of the master weights this rank owns
        self.shadow_local = self.fv.own_shadow()
        self.theta_outer = self.theta_local.detach().to(dev, copy=True)
        self.delta = torch.zeros_like(self.theta_outer)
        self.momentum_buffer: torch.Tensor | None = None
        self.offloaded_parameters = []
        if self.fv.sharded:
            # ZeRO-sharded worker: parameter boundaries do not align with the shard, the outer optimizer sees one
            # flat parameter (SGD is element-wise, so the update is identical)
            op = torch.nn.Parameter(self.theta_outer, requires_grad=True)
            op.grad = self.delta
            self.offloaded_parameters.append(op)
        else:
            for i, _ in enumerate(self.fv.params):
                op = torch.nn.Parameter(self.fv.view_of(self.theta_outer, i), requires_grad=True)
                op.grad = self.fv.view_of(self.delta, i)
                self.offloaded_parameters.append(op)
        self.optimizer: torch.optim.Optimizer = optimizer(self.offloaded_parameters)
        self.scheduler_inner_optimizer = scheduler(self.inner_optimizer) if scheduler is not None else None
        self.local_epoch = 0
        self.averaging_in_progress = False
        self.state_sharing_priority = 0
        self.custom_gradients = True
        self.offload_optimizer = True
        self.average_state_every = average_state_every
        self._adopt_sgd_momentum()

    # -- fused-kernel eligibility: torch.optim.SGD with plain (Nesterov) momentum -----------------------
    def _sgd_hparams(self):
        opt = self.optimizer
        if type(opt) is not torch.optim.SGD or len(opt.param_groups) != 1:
            return None
        g = opt.param_groups[0]
        if g.get("dampening", 0) != 0 or g.get("weight_decay", 0) != 0 or g.get("maximize", False):
            return None
        return g

    def _adopt_sgd_momentum(self) -> None:
        """Give torch.optim.SGD momentum buffers that are views of ONE flat buffer, so the fused kernel and the torch
        state_dict see the same memory.  A zero buffer is equivalent to torch's lazily created one
        (first step: buf = 0*mu + g)."""
        g = self._sgd_hparams()
        if g is None or g.get("momentum", 0) == 0:
            return
        if self.momentum_buffer is None:
            self.momentum_buffer = torch.zeros_like(self.theta_outer)
        for i, op in enumerate(self.offloaded_parameters):
            st = self.optimizer.state[op]
            view = self.momentum_buffer if self.fv.sharded else self.fv.view_of(self.momentum_buffer, i)
            if "momentum_buffer" in st and st["momentum_buffer"] is not None and st["momentum_buffer"].data_ptr() != view.data_ptr():
                view.copy_(st["momentum_buffer"])
            st["momentum_buffer"] = view

    def reload_optimizer_state(self) -> None:
        """Call after ``self.optimizer.load_state_dict`` (torch replaces state tensors with copies)."""
        self._adopt_sgd_momentum()

    # -- the outer step ---------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, *, increment_epoch: bool = True, optimizer_step: bool = True, averaging_round: bool = False,
             zero_grad: bool = False, fused_solo: bool = False, members: list[int] | None = None, **_ignored) -> None:
        """Outer optimizer step on theta_outer using the (already averaged) pseudo-gradient in its .grad buffers, then
        theta_local <- theta_outer.  ``fused_solo``: single-worker form, delta computed inside the kernel."""
        if optimizer_step:
            g = self._sgd_hparams()
            same_dev = self.theta_outer.device == self.theta_local.device
            if g is not None and same_dev and (g.get("momentum", 0) != 0):
                K.nesterov_outer(self.theta_outer, self.momentum_buffer, None if fused_solo else self.delta,
                                 self.theta_local, self.shadow_local, g["lr"], g["momentum"], bool(g.get("nesterov", False)))
                self.fv.gather_compute_weights()      # also marks derived weight copies dirty
            else:
                if fused_solo:
                    K.pseudo_grad(self.theta_outer, self.theta_local.to(self.theta_outer.device), self.delta)
                self.optimizer.step()
                self.apply_optimizer_parameters()
        if averaging_round and comm.group_size(self.dht.group if self.dht else None) > 1:
            self.average_state(members)
        if zero_grad:
            self.delta.zero_()
        if increment_epoch:
            self.local_epoch += 1

    @torch.no_grad()
    def apply_optimizer_parameters(self) -> None:
        """theta_local <- theta_outer (+ compute-dtype shadow).  hivemind: _apply_optimizer_parameters_."""
        self.theta_local.copy_(self.theta_outer, non_blocking=True)
        if self.shadow_local is not None:
            if self.shadow_local.dtype == torch.bfloat16:
                K.cast_to_bf16(self.theta_local, self.shadow_local)
            else:
                self.shadow_local.copy_(self.theta_local)
        self.fv.gather_compute_weights()

    @torch.no_grad()
    def average_state(self, members: list[int] | None = None) -> None:
        """Parameter (+momentum) averaging round - the drift-repair path (reference H2, hivemind_diloco.py:654-665).
        ``members``: restrict it to the peers of an elastic round (point-to-point transfers)."""
        group = self.dht.group if self.dht else None
        if members is not None and len(members) < comm.group_size(group):
            if len(members) > 1:
                comm.p2p_all_reduce_mean_(self.theta_outer, members, group)
                if self.momentum_buffer is not None:
                    comm.p2p_all_reduce_mean_(self.momentum_buffer, members, group)
                self.apply_optimizer_parameters()
            return
        comm.all_reduce_avg_(self.theta_outer, group)
        if self.momentum_buffer is not None:
            comm.all_reduce_avg_(self.momentum_buffer, group)
        self.apply_optimizer_parameters()

    @torch.no_grad()
    def load_state_from_peers(self, src: int = 0) -> None:
        """Download theta_outer, outer-optimizer state and the epoch from a peer (reference H3)."""
        group = self.dht.group if self.dht else None
        if comm.group_size(group) <= 1:
            return
        comm.broadcast_(self.theta_outer, src, group)
        if self.momentum_buffer is not None:
            comm.broadcast_(self.momentum_buffer, src, group)
        ep = torch.tensor([self.local_epoch], dtype=torch.int64, device=comm_device(group, torch.device("cpu")))
        comm.broadcast_(ep, src, group)
        self.local_epoch = int(ep.item())
        self.apply_optimizer_parameters()


# ====================================================================================================== the optimizer
class DiLoCoOptimizer:
    """Two-level DiLoCo optimizer with the constructor/attribute surface of the reference class
    (hivemind_diloco.py:303-738).

    :param dht: swarm handle (``opendiloco_b200.parallel.swarm.DHT``); None = this process is the only worker.
    :param batch_size: samples contributed per ``step()`` call
    :param num_inner_steps: H, inner steps per outer step
    :param outer_optimizer: factory ``params -> torch.optim.Optimizer`` (DiLoCo: SGD lr 0.7, momentum 0.9, nesterov)
    :param inner_optimizer: optimizer instance or factory ``params -> optimizer`` (DiLoCo: AdamW)
    :param scheduler: factory ``inner_optimizer -> LRScheduler`` (stepped once per inner step)
    :param grad_compression: codec from ``opendiloco_b200.parallel.compression`` applied to the pseudo-gradient round
    :param offload_device: where theta_outer lives; default = the parameters' device (HBM)
    """

    def __init__(self, *, dht: DHT | None = None, run_id: str = "diloco", batch_size: int, num_inner_steps: int,
                 outer_optimizer: Callable, inner_optimizer, params: Iterable | None = None,
                 scheduler: Callable | None = None, averager_opts: dict | None = None, grad_compression=None,
                 tracker_opts: dict | None = None, all_reduce_strategy: AllReduceStrategy = AllReduceStrategy.WAIT_FOR_ALL,
                 timeout_waiting_for_peers: float | None = None, matchmaking_time: float | None = 15.0,
                 averaging_timeout: float | None = 60.0, average_state_every: int = 1, verbose: bool = False,
                 offload_device=None, fused_collective: bool | None = None, **kwargs):
        self._check_kwargs(kwargs)
        all_reduce_strategy = AllReduceStrategy(all_reduce_strategy) if not isinstance(all_reduce_strategy, AllReduceStrategy) \
            else all_reduce_strategy
        if timeout_waiting_for_peers is not None and all_reduce_strategy == AllReduceStrategy.NO_WAIT:
            raise ValueError("You cannot use timeout_waiting_for_peers with NO_WAIT strategy, use WAIT_FOR_ALL instead")
        if timeout_waiting_for_peers is not None and matchmaking_time is not None and timeout_waiting_for_peers < matchmaking_time:
            raise ValueError("timeout_waiting_for_peers must be greater than matchmaking_time")
        if all_reduce_strategy == AllReduceStrategy.WAIT_FOR_ALL and timeout_waiting_for_peers is None:
            timeout_waiting_for_peers = DEFAULT_TIMEOUT_WAITING_FOR_PEERS
        for factory in (outer_optimizer, scheduler):
            if not (callable(factory) or factory is None):
                raise TypeError("You need to pass inner and outer optimizer as well as scheduler as callable")
        if params is None:
            raise ValueError("params is required")
        params = list(params)   # model.parameters() is a generator and two optimizers consume it

        self.dht, self.run_id = dht, run_id
        self.all_reduce_strategy, self.timeout_waiting_for_peers = all_reduce_strategy, timeout_waiting_for_peers
        self.matchmaking_time = matchmaking_time if matchmaking_time is not None else 15.0
        self.averaging_timeout = averaging_timeout
        self.num_inner_steps, self.batch_size_per_step = num_inner_steps, batch_size
        self.average_state_every = average_state_every
        self.status_loglevel = 20 if verbose else 10
        self.client_mode = self.auxiliary = False
        self.delay_optimizer_step = self.delay_grad_averaging = self.delay_state_averaging = False
        self.scheduled_diloco_grads: StepControl | None = None
        self.scheduled_state: StepControl | None = None

        if isinstance(inner_optimizer, torch.optim.Optimizer):
            self.inner_optimizer = inner_optimizer
        elif callable(inner_optimizer):
            self.inner_optimizer = inner_optimizer(params=params) if _accepts_kw(inner_optimizer) else inner_optimizer(params)
        else:
            raise TypeError(f"Expected inner_optimizer to be an Optimizer or a factory, got {type(inner_optimizer)}")

        fv = getattr(self.inner_optimizer, "fv", None)     # FusedAdamW already flattened the parameters
        self.state_averager = DiLoCoStateAverager(params=params, optimizer=outer_optimizer,
                                                  inner_optimizer=self.inner_optimizer, num_inner_steps=num_inner_steps,
                                                  scheduler=scheduler, dht=dht, prefix=f"{run_id}_state_averager",
                                                  offload_device=offload_device, flat_view=fv,
                                                  average_state_every=average_state_every, **(averager_opts or {}))
        sa = self.state_averager
        self.diloco_grad_averager = DiLoCoGradAverager(
            main_parameters=sa.main_parameters, offloaded_optimizer=sa.optimizer, dht=dht,
            prefix=f"{run_id}_grad_averager", compression=grad_compression,
            flat=(sa.theta_outer, sa.delta, sa.theta_local), min_matchmaking_time=self.matchmaking_time)
        topts = dict(tracker_opts or {})
        topts.pop("private_key", None)
        topts.setdefault("max_refresh_period", 2)
        self.tracker = DiloCoProgressTracker(batch_size, num_inner_steps, dht=dht, prefix=run_id,
                                             target_batch_size=batch_size * num_inner_steps,
                                             publish=(all_reduce_strategy == AllReduceStrategy.NO_WAIT
                                                      or getattr(dht, "board", None) is not None), **topts)
        self._schema_hash = self._compute_schema_hash()
        self._drifted = False
        self._fp_pending = None          # (event, pinned [max, -min] fingerprint of theta_outer after the last full round)
        self._ns = self._make_namespace()
        self._fused = None
        want_fused = fused_collective if fused_collective is not None else True
        if want_fused and self.num_peers > 1 and sa.theta_outer.is_cuda and (grad_compression is None or grad_compression.is_identity
                                                                              or getattr(grad_compression, "fusable", False)):
            from .fused_outer import try_make_fused_outer

            self._fused = try_make_fused_outer(self, grad_compression)
        self.last_outer_step_seconds = 0.0

    # ------------------------------------------------------------------------------------------ validation
    @staticmethod
    def _check_kwargs(kwargs: dict) -> None:
        """Same rejections as the reference (hivemind_diloco.py:408-444)."""
        if "optimizer" in kwargs:
            raise KeyError("optimizer should not be passed to DiLoCoOptimizer, pass rather to outer_optimizer")
        if "use_local_updates" in kwargs:
            if kwargs.pop("use_local_updates") is False:
                raise ValueError("You cannot use DiLoCo without local updates")
        if "offload_optimizer" in kwargs:
            if kwargs.pop("offload_optimizer") is False:
                raise ValueError("offload_optimizer=False, is not supported in DiLoCo for now")
        for name in ("delay_state_averaging", "delay_grad_averaging", "delay_optimizer_step"):
            if kwargs.pop(name, False) is True:
                raise ValueError(f"{name} is not supported in DiLoCo for now")
        if "target_batch_size" in kwargs:
            raise KeyError("DiLoCo does not have a target_batch_size, use batch_size with num_inner_steps")
        if "batch_size_per_step" in kwargs:
            raise KeyError("DiLoCo does not have a batch_size_per_step, use batch_size instead")
        # hivemind transport knobs that have no meaning on NVLink are accepted and ignored
        for name in ("state_averaging_compression", "load_state_compression", "allreduce_timeout", "next_chunk_timeout",
                     "load_state_timeout", "shutdown_timeout", "reuse_grad_buffers", "grad_averager_factory",
                     "state_averager_opts", "extra_tensors", "request_timeout", "target_group_size", "part_size_bytes"):
            kwargs.pop(name, None)
        if kwargs:
            raise TypeError(f"unexpected arguments for DiLoCoOptimizer: {sorted(kwargs)}")

    def _make_namespace(self) -> str:
        """"<incarnation nonce>/g<first global rank of the outer group>": agreed on collectively at construction."""
        group = self.dht.group if self.dht is not None else None
        if comm.group_size(group) <= 1:
            return "solo"
        dev = comm_device(group, torch.device("cpu"))
        nonce = torch.tensor([time.time_ns() & ((1 << 62) - 1)], dtype=torch.int64, device=dev)
        comm.broadcast_(nonce, 0, group)
        first = dist.get_process_group_ranks(group)[0] if group is not dist.group.WORLD else 0
        return f"{int(nonce.item()):x}/g{first}"

    # ------------------------------------------------------------------------------------------ state fingerprints
    def _post_state_fingerprint(self) -> None:
        """After a full round: wrap-around integer checksum of theta_outer (+ momentum), all-reduced to (max, -min) over
        the swarm and copied to pinned memory WITHOUT a host sync; looked at when the next round starts."""
        sa = self.state_averager
        group = self.dht.group if self.dht is not None else None
        if comm.group_size(group) <= 1 or sa.theta_outer.dtype != torch.float32:
            return
        if self._fused is not None and self._fused.fingerprint is not None and self._fused.sharded:
            fp = self._fused.fingerprint.clone()                      # produced by the fused kernel's last pass
        else:
            # one pass over theta_outer at HBM speed (a momentum that drifted shows up in theta_outer one round later)
            fp = K.checksum_i32(sa.theta_outer)
        pair = torch.stack([fp.reshape(()), -fp.reshape(())]).to(comm_device(group, fp.device))
        dist.all_reduce(pair, op=dist.ReduceOp.MAX, group=group)          # [max fp, -min fp]
        if pair.is_cuda:
            host = torch.empty(2, dtype=torch.int64, pin_memory=True)
            host.copy_(pair, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            self._fp_pending = (ev, host)
        else:
            self._fp_pending = (None, pair.clone())

    def _state_diverged(self) -> bool:
        """Did the workers hold different theta_outer / momentum after the previous full round?  (Same answer on every
        worker: it is computed from an all-reduced value.)"""
        if self._fp_pending is None:
            return False
        ev, host = self._fp_pending
        self._fp_pending = None
        if ev is not None:
            ev.synchronize()             # recorded a whole epoch ago
        diverged = int(host[0]) != -int(host[1])
        if diverged:
            logger.warning("theta_outer / outer momentum differ between workers: running a state-averaging round")
        return diverged

    def _compute_schema_hash(self) -> int:
        return hash(tuple(tuple(p.shape) for p in self.state_averager.offloaded_parameters))

    # ------------------------------------------------------------------------------------------ views
    @property
    def num_peers(self) -> int:
        return self.dht.num_peers if self.dht is not None else 1

    @property
    def local_epoch(self) -> int:
        return self.state_averager.local_epoch

    @property
    def param_groups(self):
        """Inner optimizer is the main optimizer (hivemind_diloco.py:692-695)."""
        return self.inner_optimizer.param_groups

    @property
    def state(self):
        return self.inner_optimizer.state

    # ------------------------------------------------------------------------------------------ step
    def step(self, closure: Callable | None = None, batch_size: int | None = None, scaler=None):
        """Inner step; every ``num_inner_steps`` calls also the outer step (hivemind_diloco.py:483-558).
        ``scaler``: a GradScaler applied to the INNER step only - pseudo-gradients are never scaled."""
        if scaler is not None and closure is not None:
            raise ValueError("You cannot use closure and scaler at the same time")
        batch_size = batch_size if batch_size is not None else self.batch_size_per_step
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if self._fused is not None:
            self._fused.poll_timeout(block=False)        # did the last fused outer round give up on a peer?
        if self._should_load_state_from_peers():
            logger.log(self.status_loglevel, "Peer is out of sync")
            if self.all_reduce_strategy == AllReduceStrategy.NO_WAIT:
                if self._resync_from_swarm():       # point-to-point download served by the next round's leader
                    return loss
            else:
                self.load_state_from_peers()
                return loss
        self.tracker.report_local_progress(self.local_epoch, self.tracker.local_progress.samples_accumulated + batch_size)
        self._maybe_schedule_gradient_averaging()
        if scaler is not None:
            if _is_fused(self.inner_optimizer):
                self.inner_optimizer.step_scaled(scaler)      # unscale + inf check + clip + AdamW in the fused kernels
            else:
                scaler.step(self.inner_optimizer)
            from ..utils.training import found_inf_grad

            if found_inf_grad(self.inner_optimizer, scaler):
                logger.log(self.status_loglevel, f"Found inf grad at step {self.tracker.real_step}")
        else:
            self.inner_optimizer.step()
        if self.state_averager.scheduler_inner_optimizer is not None:
            self.state_averager.scheduler_inner_optimizer.step()
        if self.tracker.ready_to_update_epoch:
            self._update_global_epoch()
        return loss

    def _should_load_state_from_peers(self) -> bool:
        return self.tracker.global_epoch > self.local_epoch + 1 and self.num_peers > 1

    def _maybe_schedule_gradient_averaging(self) -> None:
        """Pre-schedule the round when the epoch is about to end (hivemind_diloco.py:722-738)."""
        if self.num_peers <= 1:
            return
        if self.all_reduce_strategy == AllReduceStrategy.WAIT_FOR_ALL:
            eta = self.tracker.global_progress.eta_next_epoch - get_dht_time()
        else:
            eta = self.tracker.estimated_next_update_time
        if eta <= self.matchmaking_time:
            c = self.scheduled_diloco_grads
            if c is None or c.triggered or c.done():
                self.scheduled_diloco_grads = self.diloco_grad_averager.schedule_step(timeout=self.averaging_timeout)

    # ------------------------------------------------------------------------------------------ round formation
    def _key(self, *parts) -> str:
        """Board key of this swarm INCARNATION and this outer group.  The nonce (agreed on at construction) keeps a
        restarted job from replaying the round records of its previous life on a long-lived board; the group tag keeps the
        outer groups of a multi-GPU worker (one per local rank) from sharing counters and member lists."""
        return "/".join([self.run_id, self._ns, *map(str, parts)])

    def _inject_fault(self, epoch: int, me: int) -> None:
        inject = os.environ.get("ODB_FAULT_INJECT")      # "rank:epoch[:seconds]" - that worker arrives late (or never)
        if not inject:
            return
        f = inject.split(":")
        if int(f[0]) == me and int(f[1]) == epoch:
            delay = float(f[2]) if len(f) > 2 else 1e9
            logger.warning(f"fault injection: worker {me} stalls {delay:.1f}s before outer step {epoch}")
            time.sleep(min(delay, 3600.0))

    def _form_round(self) -> tuple[list[int], bool, bool] | None:
        """Matchmaking through the membership board, for both strategies (hivemind: the round is formed by whoever shows
        up inside the window; the others form the next round among themselves).  The first arrival of round r of this
        epoch is its leader: it waits until everybody still missing has arrived or the window closes -
        ``timeout_waiting_for_peers`` for WAIT_FOR_ALL ("Timeout waiting for peers, going to skip slowest peers",
        hivemind_diloco.py:584-607), ``matchmaking_time`` for NO_WAIT - then publishes the member list.  A partial round
        averages over the members only (point-to-point or a member-masked NVLink window), so workers that are still
        training - or dead - are not needed.
        Returns (sorted member ranks, am_i_leader, repair) or None when there is no board.  ``repair``: some member saw a
        split epoch or adopted the swarm state since the last full round -> this full round also averages the state.
        Store errors propagate: a broken board must not be mistaken for "everyone arrived"."""
        store = self.dht.store() if self.dht is not None else None
        n = self.num_peers
        if store is None or n <= 1:
            return None
        if self.all_reduce_strategy == AllReduceStrategy.WAIT_FOR_ALL and self.timeout_waiting_for_peers is None:
            return None                                  # handshake switched off (static swarm; kernel-time benchmarks)
        epoch, me = self.local_epoch, self.dht.rank_in_group
        self._inject_fault(epoch, me)
        window = float(self.timeout_waiting_for_peers if self.all_reduce_strategy == AllReduceStrategy.WAIT_FOR_ALL
                       else self.matchmaking_time)
        done_before = 0                                  # peers that already finished this epoch in earlier rounds
        r = 0
        mark = "D" if self._drifted else "1"
        while True:
            mkey = self._key("round", epoch, r, "members")
            if store.check([mkey]):                      # round r is closed: try the next one
                done_before += len(store.get(mkey).decode().split("|")[0].split(","))
                r += 1
                continue
            order = int(store.add(self._key("round", epoch, r, "count"), 1))
            store.set(self._key("round", epoch, r, "arrive", me), mark)
            keys = [self._key("round", epoch, r, "arrive", q) for q in range(n)]
            t_start = time.perf_counter()
            if order == 1:
                deadline = t_start + window
                everyone = False
                next_liveness = t_start + 0.5
                while time.perf_counter() < deadline:
                    if done_before == 0 and store.check(keys):       # ONE round trip when the whole swarm is punctual
                        everyone = True
                        break
                    if done_before > 0 and sum(store.check([k]) for k in keys) >= n - done_before:
                        break
                    # workers of a healthy swarm arrive within microseconds of each other: spin first, back off later
                    if time.perf_counter() - t_start > 0.002:
                        time.sleep(0.0005)
                    # failure detection: with a heartbeat board a peer whose heartbeat expired cannot arrive any more - do
                    # not sit out the whole window for it epoch after epoch (checked twice a second)
                    if time.perf_counter() >= next_liveness:
                        next_liveness = time.perf_counter() + 0.5
                        if self._nobody_else_can_come(store, keys):
                            break
                members = list(range(n)) if everyone else [q for q in range(n) if store.check([keys[q]])]
                mkeys = [keys[q] for q in members]
                marks = store.multi_get(mkeys) if hasattr(store, "multi_get") else [store.get(k) for k in mkeys]
                repair = any(bytes(x).decode() == "D" for x in marks)
                store.set(mkey, ",".join(str(q) for q in members) + "|" + ("1" if repair else "0"))
                if len(members) < n - done_before:
                    logger.log(self.status_loglevel, f"Timeout waiting for peers, going to skip slowest peers; present={members}")
                self._gc_due = epoch - 2                 # old round records are deleted after the round has been launched
                return members, True, repair
            deadline = t_start + window + float(self.averaging_timeout or 60.0)
            while not store.check([mkey]):
                if time.perf_counter() > deadline:
                    raise TimeoutError(f"round {r} of outer step {epoch}: the leader never published the member list")
                if time.perf_counter() - t_start > 0.002:
                    time.sleep(0.0005)
            rec = store.get(mkey).decode().split("|")
            members = [int(q) for q in rec[0].split(",")]
            if me in members:
                return members, False, rec[1] == "1"
            done_before += len(members)                  # my arrival raced with the leader's snapshot: next round
            r += 1

    def _nobody_else_can_come(self, store, keys: list[str]) -> bool:
        """True when every peer with a LIVE heartbeat on the membership board has arrived (only meaningful with the native
        board, ``csrc/host/rendezvous.cc``: the c10d store has no liveness information and the answer is then False)."""
        if getattr(self.dht, "board", None) is None:
            return False
        try:
            alive = {int(p.rsplit("-", 1)[1]) for p in self.dht.alive_peers()}
        except Exception:
            return False
        missing = [q for q in range(len(keys)) if not store.check([keys[q]])]
        dead = [q for q in missing if q not in alive]
        if missing and len(dead) == len(missing):
            logger.log(self.status_loglevel, f"workers {dead} have no live heartbeat: closing the round without them")
            return True
        return False

    def _gc_round_keys(self, store, epoch: int, n: int) -> None:
        """Drop the records of a long-closed epoch (best effort: c10d TCPStore and the native board can delete)."""
        if epoch < 0 or not hasattr(store, "delete_key"):
            return
        try:
            for r in range(n):
                if not store.check([self._key("round", epoch, r, "count")]):
                    break
                for k in ("members", "count"):
                    store.delete_key(self._key("round", epoch, r, k))
                for q in range(n):
                    store.delete_key(self._key("round", epoch, r, "arrive", q))
        except Exception:
            pass

    def _serve_resync_requests(self) -> None:
        """Round leader, after its outer update: hand theta_outer / momentum / epoch to peers that fell behind and asked
        for the swarm state (the reference's load_state_from_peers, served between two rounds instead of by a
        background thread)."""
        store = self.dht.store() if self.dht is not None else None
        if store is None:
            return
        sa, me = self.state_averager, self.dht.rank_in_group
        self._sync_outer_state()
        for q in range(self.num_peers):
            rkey = self._key("resync", "request", q)
            if q == me or not store.check([rkey]):
                continue
            req = store.get(rkey).decode()
            if req == "" or int(store.add(self._key("resync", "claim", q, req), 1)) != 1:
                continue                                 # nothing pending, or another leader serves it
            logger.log(self.status_loglevel, f"serving swarm state (epoch {self.local_epoch}) to worker {q}")
            store.set(self._key("resync", "grant", q, req), f"{me},{self.local_epoch}")
            bufs = [sa.theta_outer] + ([sa.momentum_buffer] if sa.momentum_buffer is not None else [])
            comm.p2p_send_(bufs, q, self.dht.group)

    def _resync_from_swarm(self) -> bool:
        """Lagging peer: ask for the swarm state on the board and block until a round leader sends it."""
        store = self.dht.store() if self.dht is not None else None
        if store is None:
            return False
        sa, me = self.state_averager, self.dht.rank_in_group
        self._resync_seq = getattr(self, "_resync_seq", 0) + 1
        req = str(self._resync_seq)
        store.set(self._key("resync", "request", me), req)
        gkey = self._key("resync", "grant", me, req)
        deadline = time.perf_counter() + float(self.averaging_timeout or 600.0)
        while not store.check([gkey]):
            if time.perf_counter() > deadline:
                store.set(self._key("resync", "request", me), "")
                logger.warning("no round leader served the state request in time; continuing with the local state")
                return False
            time.sleep(0.005)
        src, epoch = (int(x) for x in store.get(gkey).decode().split(","))
        store.set(self._key("resync", "request", me), "")
        with self.tracker.pause_updates():
            bufs = [sa.theta_outer] + ([sa.momentum_buffer] if sa.momentum_buffer is not None else [])
            comm.p2p_recv_(bufs, src, self.dht.group)
            sa.local_epoch = epoch
            sa.apply_optimizer_parameters()
            self.tracker.update_epoch(epoch)
            self.tracker.report_local_progress(epoch, samples_accumulated=0)
        # the state came from ONE peer of a swarm that split: have the next full round average it (the flag travels with
        # this worker's arrival record, so every member of that round takes the same decision)
        self._drifted = True
        logger.log(self.status_loglevel, f"adopted the swarm state of worker {src} at epoch {epoch}")
        return True

    def _update_global_epoch(self) -> None:
        """The outer step (hivemind_diloco.py:570-679)."""
        assert self._schema_hash == self._compute_schema_hash(), "parameters changed during iteration"
        t_start = time.perf_counter()
        sa, ga = self.state_averager, self.diloco_grad_averager
        members, leader, repair = None, False, False
        if self._fused is not None:
            self._fused.poll_timeout(block=True)         # the previous fused round must have completed on every rank
        if self.num_peers > 1:
            formed = self._form_round()
            if formed is not None:
                members, leader, repair = formed
                if len(members) == self.num_peers:
                    members = None                       # everybody made it: the ordinary full-group round
                else:
                    logger.log(self.status_loglevel, f"outer step {self.local_epoch}: elastic round with workers {members} "
                                                     f"of {self.num_peers}")
        self.last_round_members = members if members is not None else list(range(self.num_peers))
        with self.tracker.pause_updates():
            next_epoch = max(self.local_epoch + 1, self.tracker.global_epoch) if members is None else self.local_epoch + 1
            # State averaging (hivemind average_state_every, default 1 = every epoch; hivemind_diloco.py:630-637).  After a
            # full round every worker holds bit-identical theta_outer / momentum, so averaging them is the identity: the
            # round is therefore run when it can change something - the swarm split since the last full round
            # (``repair``, decided by the round leader from the members' arrival records, hence the same on every
            # member), the fingerprints of the previous round disagreed (``_state_diverged``), or unconditionally with
            # ODB_FORCE_STATE_AVERAGING=1.
            due = self.num_peers > 1 and self.average_state_every > 0 and next_epoch % self.average_state_every == 0
            average_state = due and members is None and (self._state_diverged() or os.environ.get("ODB_FORCE_STATE_AVERAGING") == "1")
            if members is not None:
                self._drifted = True                     # the swarm split this epoch: theta_outer differs between rounds
            elif self.num_peers > 1:
                average_state = average_state or repair
                self._drifted = False
            if members is not None or average_state:
                self._sync_outer_state()                 # these paths read / average the FULL outer momentum
            if members is not None and len(members) == 1:
                # nobody else showed up in time: this worker's own pseudo-gradient is the round
                sa.step(increment_epoch=True, optimizer_step=True, averaging_round=False, fused_solo=True)
            elif members is not None and self._fused is not None and _lib_has("odb_fused_outer_subset"):
                # partial round on the NVLink window: member-masked peer loads / stores, absent workers are not touched
                self._fused.outer_step_subset(members, self.local_epoch)
                sa.step(increment_epoch=True, optimizer_step=False, averaging_round=False)
                self._fused.poll_timeout(block=False)
            elif members is not None:
                ga.step(wait=True, timeout=self.averaging_timeout, control=self.scheduled_diloco_grads, members=members)
                ga.notify_used_averaged_gradients()
                self.scheduled_diloco_grads = None
                sa.step(increment_epoch=True, optimizer_step=True, averaging_round=average_state, members=members)
            elif self._fused is not None:
                # pseudo-grad + NVLink reduce + Nesterov in ONE kernel (replicated update when the state is about to be averaged)
                self._fused.outer_step(self.local_epoch, replicated=average_state)
                sa.step(increment_epoch=True, optimizer_step=False, averaging_round=average_state)
                self._fused.poll_timeout(block=False)
            elif self.num_peers > 1:
                logger.log(self.status_loglevel, f"Beginning optimizer step #{self.local_epoch}")
                ga.step(wait=True, timeout=self.averaging_timeout, control=self.scheduled_diloco_grads)
                logger.log(self.status_loglevel, f"Time taken for gradient all reduce: {ga.last_allreduce_seconds} sec")
                ga.notify_used_averaged_gradients()
                self.scheduled_diloco_grads = None
                sa.step(increment_epoch=True, optimizer_step=True, averaging_round=average_state)
            else:
                sa.step(increment_epoch=True, optimizer_step=True, averaging_round=False, fused_solo=True)
            if self.num_peers > 1 and members is None and self.average_state_every > 0:
                self._post_state_fingerprint()
            if self._fused is not None:
                self._fused.start_momentum_regather()    # background: owners publish their momentum slabs (side stream)
            if leader:
                self._serve_resync_requests()
                if getattr(self, "_gc_due", -1) >= 0:
                    self._gc_round_keys(self.dht.store(), self._gc_due, self.num_peers)
                    self._gc_due = -1
            if self.scheduled_state is not None and not self.scheduled_state.done():
                self.scheduled_state.cancel()
            self.scheduled_state = None
            self.tracker.update_epoch(new_epoch=sa.local_epoch)
            sa.state_sharing_priority = self.local_epoch
            logger.log(self.status_loglevel, f"Transitioning to epoch {self.local_epoch}")
        self.last_outer_step_seconds = time.perf_counter() - t_start

    def _sync_outer_state(self) -> None:
        """The sharded fused outer step keeps each momentum slab current on its owner only and re-replicates it in the
        background; anything that reads the whole outer-optimizer state (checkpoints, partial rounds, state averaging,
        serving a lagging peer) first waits for that all-gather."""
        if self._fused is not None:
            self._fused.wait_momentum()

    def update_main_param_after_outer_step(self) -> None:
        """No-op kept for API parity: the outer kernel already wrote theta_local (SURVEY.md §2.7 first quirk)."""

    # ------------------------------------------------------------------------------------------ housekeeping
    def zero_grad(self, set_to_none: bool = False) -> None:
        self.inner_optimizer.zero_grad(set_to_none=False) if _is_fused(self.inner_optimizer) else \
            self.inner_optimizer.zero_grad(set_to_none=set_to_none)

    def load_state_from_peers(self, **kwargs) -> None:
        """Adopt theta_outer / outer state / epoch from the swarm (collective over the outer group; reference:
        train_fsdp.py:348-349, hivemind_diloco.py:528-531)."""
        if self.scheduled_diloco_grads is not None:
            self.scheduled_diloco_grads.cancel()
            self.scheduled_diloco_grads = None
        self._sync_outer_state()
        with self.tracker.pause_updates():
            self.state_averager.load_state_from_peers()
            self.tracker.report_local_progress(self.local_epoch, samples_accumulated=0)

    def state_dict(self) -> dict:
        """{"state_dict_outer": outer sd (+ ["state"]["local_epoch"]), "state_dict_inner": inner sd}
        (hivemind_diloco.py:697-707) plus what the reference forgets to save: theta_outer itself and the inner-step
        phase."""
        self._sync_outer_state()
        sd_outer = self.state_averager.optimizer.state_dict()
        sd_outer["state"]["local_epoch"] = self.local_epoch
        return {
            "state_dict_outer": sd_outer,
            "state_dict_inner": self.inner_optimizer.state_dict(),
            "theta_outer": self.state_averager.theta_outer.detach().cpu().clone(),
            "samples_accumulated": self.tracker.local_progress.samples_accumulated,
            "drifted": bool(self._drifted),
        }

    def load_state_dict(self, state_dict: dict) -> None:
        sd_outer = dict(state_dict["state_dict_outer"])
        sd_outer["state"] = dict(sd_outer["state"])
        if "local_epoch" in sd_outer["state"]:
            self.state_averager.local_epoch = int(sd_outer["state"].pop("local_epoch"))
        self.state_averager.optimizer.load_state_dict(sd_outer)
        self.state_averager.reload_optimizer_state()
        self.inner_optimizer.load_state_dict(state_dict["state_dict_inner"])
        if state_dict.get("theta_outer") is not None:
            self.state_averager.theta_outer.copy_(state_dict["theta_outer"])
        self._drifted = bool(state_dict.get("drifted", False))
        self.tracker.update_epoch(self.local_epoch)
        self.tracker.report_local_progress(self.local_epoch, int(state_dict.get("samples_accumulated", 0)))

    def shutdown(self) -> None:
        self.diloco_grad_averager.shutdown()
        if self._fused is not None:
            self._fused.close()


def _lib_has(symbol: str) -> bool:
    from .. import _lib

    return _lib.has_symbol(symbol)


def _accepts_kw(fn) -> bool:
    import inspect

    try:
        sig = inspect.signature(fn)
    except (TypeError, ValueError):
        return True
    return "params" in sig.parameters or any(p.kind == p.VAR_KEYWORD for p in sig.parameters.values())


def _is_fused(opt) -> bool:
    from ..optim.fused import FusedAdamW

    return isinstance(opt, FusedAdamW)
