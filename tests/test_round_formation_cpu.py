"""Round formation on the membership board (DiLoCoOptimizer._form_round) with an in-memory store and one thread per worker:
no process group, no sleeping on real collectives - only the matchmaking protocol.

Reference behaviour being matched: hivemind matchmaking + hivemind_diloco.py:578-608 (WAIT_FOR_ALL waits up to
``timeout_waiting_for_peers`` and then "skips the slowest peers"; NO_WAIT gives late peers ``matchmaking_time``)."""
import threading
import time
from functools import partial

import torch

from opendiloco_b200.parallel.diloco import AllReduceStrategy, DiLoCoOptimizer


class MemStore:
    """Thread-safe stand-in for the c10d TCPStore / native board (set, get, check, add, multi_get, delete_key)."""

    def __init__(self):
        self.d, self.mu, self.ops = {}, threading.Lock(), 0

    def set(self, k, v):
        with self.mu:
            self.ops += 1
            self.d[k] = v.encode() if isinstance(v, str) else v

    def get(self, k):
        with self.mu:
            self.ops += 1
            return self.d[k]

    def multi_get(self, keys):
        with self.mu:
            self.ops += 1
            return [self.d[k] for k in keys]

    def check(self, keys):
        with self.mu:
            self.ops += 1
            return all(k in self.d for k in keys)

    def add(self, k, delta):
        with self.mu:
            self.ops += 1
            cur = int(self.d.get(k, b"0")) + delta
            self.d[k] = str(cur).encode()
            return cur

    def delete_key(self, k):
        with self.mu:
            self.ops += 1
            return self.d.pop(k, None) is not None


class FakeDHT:
    group = None
    board = None

    def __init__(self, store, rank, n):
        self._s, self.rank_in_group, self.num_peers, self.peer_id = store, rank, n, f"worker-{rank}"

    def store(self):
        return self._s

    def peer_ids(self):
        return [f"worker-{r}" for r in range(self.num_peers)]


def make_worker(store, rank, n, strategy, **kw):
    p = torch.nn.Parameter(torch.zeros(8))
    opt = DiLoCoOptimizer(dht=None, run_id="t", batch_size=1, num_inner_steps=1, params=[p],
                          outer_optimizer=partial(torch.optim.SGD, lr=1.0), inner_optimizer=partial(torch.optim.SGD, lr=1.0),
                          all_reduce_strategy=strategy, **kw)
    opt.dht = FakeDHT(store, rank, n)        # the ctor ran solo (no process group); the protocol only needs the board
    opt._ns = "test/g0"
    return opt


def run_round(workers, delays):
    out = [None] * len(workers)

    def go(i):
        time.sleep(delays[i])
        out[i] = workers[i]._form_round()

    th = [threading.Thread(target=go, args=(i,)) for i in range(len(workers))]
    [t.start() for t in th]
    [t.join(30) for t in th]
    return out


def test_punctual_swarm_forms_one_full_round_with_one_leader():
    store = MemStore()
    ws = [make_worker(store, r, 4, AllReduceStrategy.WAIT_FOR_ALL, timeout_waiting_for_peers=5.0, matchmaking_time=1.0) for r in range(4)]
    t0 = time.perf_counter()
    res = run_round(ws, [0.0, 0.01, 0.02, 0.03])
    assert time.perf_counter() - t0 < 2.0                      # nobody sat out the 5 s window
    assert all(r[0] == [0, 1, 2, 3] for r in res)
    assert sum(r[1] for r in res) == 1                          # exactly one leader
    assert not any(r[2] for r in res)                           # nobody drifted: no repair


def test_wait_for_all_skips_the_slowest_peer_after_the_timeout():
    store = MemStore()
    ws = [make_worker(store, r, 3, AllReduceStrategy.WAIT_FOR_ALL, timeout_waiting_for_peers=0.6, matchmaking_time=0.2) for r in range(3)]
    res = run_round(ws, [0.0, 0.05, 1.5])                       # worker 2 arrives after the window
    assert res[0][0] == res[1][0] == [0, 1]
    assert res[2][0] == [2] and res[2][1]                       # the straggler leads (and is alone in) round 1 of the epoch
    assert [r[1] for r in res[:2]].count(True) == 1


def test_drift_mark_of_one_member_becomes_the_repair_decision_of_all():
    store = MemStore()
    ws = [make_worker(store, r, 3, AllReduceStrategy.NO_WAIT, matchmaking_time=1.0) for r in range(3)]
    ws[2]._drifted = True                                       # e.g. it adopted a peer's state, or resumed from a checkpoint
    res = run_round(ws, [0.0, 0.02, 0.04])
    assert all(r[0] == [0, 1, 2] and r[2] for r in res)         # every member is told to run the state-averaging round


def test_keys_are_namespaced_per_incarnation_and_group():
    """A restarted job (new nonce) or another outer group of the same worker must not see this group's round records."""
    store = MemStore()
    a = [make_worker(store, r, 2, AllReduceStrategy.WAIT_FOR_ALL, timeout_waiting_for_peers=2.0, matchmaking_time=0.5) for r in range(2)]
    assert all(r[0] == [0, 1] for r in run_round(a, [0.0, 0.01]))
    b = [make_worker(store, r, 2, AllReduceStrategy.WAIT_FOR_ALL, timeout_waiting_for_peers=2.0, matchmaking_time=0.5) for r in range(2)]
    for w in b:
        w._ns = "test/g1"                                       # same epoch 0, other namespace: starts from a clean slate
    res = run_round(b, [0.0, 0.01])
    assert all(r[0] == [0, 1] for r in res) and sum(r[1] for r in res) == 1
    assert any(k.startswith("t/test/g0/") for k in store.d) and any(k.startswith("t/test/g1/") for k in store.d)


def test_leader_does_not_wait_the_window_for_a_peer_without_heartbeat():
    """With a heartbeat board (csrc/host/rendezvous.cc) a dead worker is skipped as soon as its heartbeat has expired instead
    of after ``timeout_waiting_for_peers`` every epoch (the reference learns the same from expired DHT records)."""
    store = MemStore()
    ws = [make_worker(store, r, 3, AllReduceStrategy.WAIT_FOR_ALL, timeout_waiting_for_peers=30.0, matchmaking_time=1.0) for r in range(2)]
    for w in ws:
        w.dht.board = object()                                  # "a board is attached"
        w.dht.alive_peers = lambda: ["worker-0", "worker-1"]    # worker-2's heartbeat has expired
    t0 = time.perf_counter()
    res = run_round(ws, [0.0, 0.02])
    assert time.perf_counter() - t0 < 3.0                       # not the 30 s window
    assert res[0][0] == res[1][0] == [0, 1]


def test_handshake_can_be_switched_off():
    store = MemStore()
    w = make_worker(store, 0, 2, AllReduceStrategy.WAIT_FOR_ALL, timeout_waiting_for_peers=2.0, matchmaking_time=0.5)
    w.timeout_waiting_for_peers = None
    assert w._form_round() is None and store.ops == 0
