"""Tracing / timing helpers the reference lacks entirely: NVTX ranges around the phases of a step and
CUDA-event timers that never force a host sync on the hot path (times are resolved lazily, one step later).

    with nvtx_range("outer_step"): ...
    timer = StepTimer(); timer.start("inner"); ...; timer.stop("inner"); timer.summary() -> {"inner_ms": ...}

`ncu` / sanitizer recipes live in profiles/ (profile_step.py, summarize_launches.py, top_stalls.py, run_sanitizer.sh)."""
from __future__ import annotations

import contextlib
import os

import torch

_NVTX = os.environ.get("ODB_NVTX", "0") == "1"


@contextlib.contextmanager
def nvtx_range(name: str):
    if _NVTX and torch.cuda.is_available():
        torch.cuda.nvtx.range_push(name)
        try:
            yield
        finally:
            torch.cuda.nvtx.range_pop()
    else:
        yield


class StepTimer:
    """Named CUDA-event stopwatches; `summary()` reads events recorded earlier (waits only if they are not done yet)."""

    def __init__(self, enabled: bool = True):
        self.enabled = enabled and torch.cuda.is_available()
        self._open: dict[str, torch.cuda.Event] = {}
        self._pairs: dict[str, list] = {}

    def start(self, name: str) -> None:
        if self.enabled:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self._open[name] = ev

    def stop(self, name: str) -> None:
        if self.enabled and name in self._open:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self._pairs.setdefault(name, []).append((self._open.pop(name), ev))

    def summary(self, reset: bool = True) -> dict[str, float]:
        out = {}
        for name, pairs in self._pairs.items():
            done = [(a, b) for a, b in pairs if b.query()]
            if done:
                out[f"{name}_ms"] = sum(a.elapsed_time(b) for a, b in done) / len(done)
            if reset:
                self._pairs[name] = [(a, b) for a, b in pairs if not b.query()]
        return out
