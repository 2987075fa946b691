"""Python face of the native rendezvous / membership board (``csrc/host/rendezvous.cc``).

The reference's workers find each other through hivemind's DHT, carried by the Go ``p2pd`` daemon (SURVEY.md §2.2 E11,
``run_training.sh:45``).  Here one process hosts a small TCP service (``RendezvousServer``) and every worker, launcher or
monitor talks to it through ``RendezvousClient``: key-value records (progress), atomic counters (arrivals of an outer
step) and heartbeats with expiry (who is alive).  It does not need ``torch.distributed`` - workers of different torchrun
jobs and tools outside the job can share one board.  ``as_store()`` gives the c10d-store surface (set / get / check /
add) that ``DiloCoProgressTracker`` and the arrival handshake use.

    server = RendezvousServer(port=29400)                       # worker 0 / launcher
    board = RendezvousClient("127.0.0.1", 29400, peer_id="worker-3")
    board.heartbeat(ttl=30.0); board.alive_peers()            # -> ["worker-0", ..., "worker-3"]
"""
from __future__ import annotations

import ctypes
import threading
import time

from .. import _lib

c_void_p, c_int, c_ll, c_char_p = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_char_p
_bound = False


def _host():
    global _bound
    lib = _lib.host_lib()
    if lib is None or not hasattr(lib, "odb_rdv_server_start"):
        raise RuntimeError("libodbhost.so with the rendezvous service is not built: run `python -m opendiloco_b200.build`")
    if not _bound:
        lib.odb_rdv_server_start.argtypes, lib.odb_rdv_server_start.restype = [c_int, ctypes.POINTER(c_int)], c_void_p
        lib.odb_rdv_server_stop.argtypes, lib.odb_rdv_server_stop.restype = [c_void_p], None
        lib.odb_rdv_server_num_keys.argtypes, lib.odb_rdv_server_num_keys.restype = [c_void_p], c_ll
        lib.odb_rdv_connect.argtypes, lib.odb_rdv_connect.restype = [c_char_p, c_int, c_int], c_void_p
        lib.odb_rdv_close.argtypes, lib.odb_rdv_close.restype = [c_void_p], None
        lib.odb_rdv_set.argtypes, lib.odb_rdv_set.restype = [c_void_p, c_char_p, c_int, c_char_p, c_ll], c_int
        lib.odb_rdv_get.argtypes, lib.odb_rdv_get.restype = [c_void_p, c_char_p, c_int, c_void_p, c_ll], c_ll
        lib.odb_rdv_add.argtypes, lib.odb_rdv_add.restype = [c_void_p, c_char_p, c_int, c_ll, ctypes.POINTER(c_ll)], c_int
        lib.odb_rdv_del.argtypes, lib.odb_rdv_del.restype = [c_void_p, c_char_p, c_int], c_int
        lib.odb_rdv_count.argtypes, lib.odb_rdv_count.restype = [c_void_p, c_char_p, c_int], c_ll
        lib.odb_rdv_beat.argtypes, lib.odb_rdv_beat.restype = [c_void_p, c_char_p, c_int, c_int], c_int
        lib.odb_rdv_peers.argtypes, lib.odb_rdv_peers.restype = [c_void_p, c_void_p, c_ll], c_ll
        _bound = True
    return lib


def available() -> bool:
    lib = _lib.host_lib()
    return lib is not None and hasattr(lib, "odb_rdv_server_start")


class RendezvousServer:
    """Hosts the board on ``port`` (0 = pick a free one; see ``.port``) until ``stop()`` / garbage collection."""

    def __init__(self, port: int = 0):
        bound = c_int(0)
        self._h = _host().odb_rdv_server_start(int(port), ctypes.byref(bound))
        if not self._h:
            raise OSError(f"rendezvous server could not listen on port {port}")
        self.port = int(bound.value)

    def num_keys(self) -> int:
        return int(_host().odb_rdv_server_num_keys(self._h))

    def stop(self) -> None:
        if self._h:
            _host().odb_rdv_server_stop(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.stop()
        except Exception:
            pass


def _b(x) -> bytes:
    return x if isinstance(x, (bytes, bytearray)) else str(x).encode()


class RendezvousClient:
    def __init__(self, host: str = "127.0.0.1", port: int = 29400, peer_id: str | None = None, connect_timeout: float = 30.0):
        self._lib = _host()
        self._h = self._lib.odb_rdv_connect(host.encode(), int(port), int(connect_timeout * 1000))
        if not self._h:
            raise ConnectionError(f"no rendezvous server at {host}:{port} after {connect_timeout:.0f}s")
        self.host, self.port, self.peer_id = host, int(port), peer_id
        self._beat_thread: threading.Thread | None = None
        self._beat_stop = threading.Event()

    # -- key-value records ------------------------------------------------------------------------
    def set(self, key, value) -> None:
        k, v = _b(key), _b(value)
        if self._lib.odb_rdv_set(self._h, k, len(k), v, len(v)) != 0:
            raise ConnectionError("rendezvous set failed")

    def get(self, key, default=None):
        k = _b(key)
        cap = 1 << 12
        while True:
            buf = ctypes.create_string_buffer(cap)
            n = self._lib.odb_rdv_get(self._h, k, len(k), buf, cap)
            if n >= 0:
                return buf.raw[:n]
            if n == -1:
                return default
            if n == -2:
                raise ConnectionError("rendezvous get failed")
            cap = int(-n - 2)                      # the value is larger than the buffer: retry with the reported size

    def delete(self, key) -> bool:
        k = _b(key)
        return self._lib.odb_rdv_del(self._h, k, len(k)) == 0

    def add(self, key, delta: int = 1) -> int:
        """Atomic fetch-add on a shared int64 counter (created at 0); returns the new value."""
        k, out = _b(key), c_ll(0)
        if self._lib.odb_rdv_add(self._h, k, len(k), int(delta), ctypes.byref(out)) != 0:
            raise ConnectionError("rendezvous add failed")
        return int(out.value)

    def count(self, prefix) -> int:
        p = _b(prefix)
        n = self._lib.odb_rdv_count(self._h, p, len(p))
        if n < 0:
            raise ConnectionError("rendezvous count failed")
        return int(n)

    def wait(self, keys, timeout: float) -> bool:
        """True once every key exists; False after ``timeout`` seconds (the arrival handshake of an outer step)."""
        deadline = time.perf_counter() + timeout
        keys = [_b(k) for k in keys]
        while True:
            if all(self.get(k) is not None for k in keys):
                return True
            if time.perf_counter() >= deadline:
                return False
            time.sleep(0.001)

    # -- liveness -----------------------------------------------------------------------------------
    def heartbeat(self, ttl: float = 30.0, peer_id: str | None = None) -> None:
        pid = _b(peer_id or self.peer_id or "anonymous")
        if self._lib.odb_rdv_beat(self._h, pid, len(pid), int(ttl * 1000)) != 0:
            raise ConnectionError("rendezvous heartbeat failed")

    def alive_peers(self) -> list[str]:
        cap = 1 << 14
        while True:
            buf = ctypes.create_string_buffer(cap)
            n = self._lib.odb_rdv_peers(self._h, buf, cap)
            if n >= 0:
                return [p for p in buf.raw[:n].decode().split("\n") if p]
            if n == -2:
                raise ConnectionError("rendezvous peers failed")
            cap = int(-n - 2)

    def start_heartbeat(self, ttl: float = 30.0, period: float | None = None) -> None:
        """Background thread that keeps this peer alive on the board (period defaults to ttl / 3)."""
        if self._beat_thread is not None:
            return
        period = period if period is not None else ttl / 3.0
        self.heartbeat(ttl)

        def loop():
            while not self._beat_stop.wait(period):
                try:
                    self.heartbeat(ttl)
                except Exception:
                    return

        self._beat_thread = threading.Thread(target=loop, name="odb-heartbeat", daemon=True)
        self._beat_thread.start()

    # -- c10d-store surface -------------------------------------------------------------------------
    def as_store(self) -> "BoardStore":
        return BoardStore(self)

    def close(self) -> None:
        self._beat_stop.set()
        if self._beat_thread is not None:
            self._beat_thread.join(timeout=2.0)
            self._beat_thread = None
        if self._h:
            self._lib.odb_rdv_close(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


class BoardStore:
    """The subset of ``torch.distributed.Store`` the swarm code uses, backed by the native board."""

    def __init__(self, client: RendezvousClient):
        self.client = client

    def set(self, key: str, value) -> None:
        self.client.set(key, value)

    def get(self, key: str) -> bytes:
        v = self.client.get(key)
        if v is None:
            raise KeyError(key)
        return v

    def check(self, keys) -> bool:
        return all(self.client.get(k) is not None for k in keys)

    def add(self, key: str, delta: int) -> int:
        return self.client.add(key, delta)

    def delete_key(self, key: str) -> bool:
        return self.client.delete(key)


def parse_address(addr: str) -> tuple[str, int] | None:
    """``odb://host:port`` (or ``tcp://host:port``) -> (host, port); anything else -> None."""
    for scheme in ("odb://", "tcp://"):
        if addr.startswith(scheme):
            host, _, port = addr[len(scheme):].rpartition(":")
            if host and port.isdigit():
                return host, int(port)
    return None


if __name__ == "__main__":          # python -m opendiloco_b200.parallel.rendezvous --serve 29400
    import argparse
    import signal

    ap = argparse.ArgumentParser(description="host the swarm membership board")
    ap.add_argument("--serve", type=int, default=29400, metavar="PORT")
    a = ap.parse_args()
    srv = RendezvousServer(a.serve)
    print(f"swarm board listening on odb://0.0.0.0:{srv.port}  (export ODB_BOARD=odb://<host>:{srv.port})", flush=True)
    stop = threading.Event()
    signal.signal(signal.SIGTERM, lambda *_: stop.set())
    signal.signal(signal.SIGINT, lambda *_: stop.set())
    stop.wait()
    srv.stop()
