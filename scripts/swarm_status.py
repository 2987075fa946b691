#!/usr/bin/env python
"""Watch a running swarm from outside the job through the native membership board (``csrc/host/rendezvous.cc``).

    python scripts/swarm_status.py odb://127.0.0.1:29400 [--run-id llama] [--galaxy-size 8] [--watch 5]

Prints who is alive (heartbeats that have not expired) and every worker's last progress record (epoch, samples of the
current epoch, samples/s, age) - what hivemind's DHT lets any peer see about the others (hivemind_diloco.py:269-272)."""
from __future__ import annotations

import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from opendiloco_b200.parallel import rendezvous as rdv  # noqa: E402


def snapshot(board: rdv.RendezvousClient, prefix: str, n: int) -> list[str]:
    now = time.time()
    alive = set(board.alive_peers())
    lines = [f"alive: {sorted(alive) or '-'}"]
    for r in range(n):
        pid = f"worker-{r}"
        rec = board.get(f"{prefix}_progress/{pid}")
        if rec is None:
            lines.append(f"  {pid}: no record{'' if pid not in alive else ' (alive)'}")
            continue
        epoch, samples, sps, t = rec.decode().split(",")
        lines.append(f"  {pid}: epoch {epoch}  samples {samples}  {float(sps):9.1f} samples/s  last report {now - float(t):6.1f} s ago"
                     f"{'' if pid in alive else '  [no heartbeat]'}")
    return lines


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("address", help="odb://host:port of the board (ODB_BOARD of the job)")
    ap.add_argument("--prefix", default="llama", help="tracker prefix = the run_id of the job (train_fsdp uses \"llama\")")
    ap.add_argument("--galaxy-size", type=int, default=8)
    ap.add_argument("--watch", type=float, default=0.0, help="refresh period in seconds (0 = print once)")
    a = ap.parse_args()
    hp = rdv.parse_address(a.address)
    if hp is None:
        raise SystemExit(f"bad address {a.address!r} (expected odb://host:port)")
    board = rdv.RendezvousClient(hp[0], hp[1], connect_timeout=5.0)
    try:
        while True:
            print("\n".join(snapshot(board, a.prefix, a.galaxy_size)), flush=True)
            if a.watch <= 0:
                break
            time.sleep(a.watch)
    finally:
        board.close()


if __name__ == "__main__":
    main()
