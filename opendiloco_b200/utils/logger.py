"""Text logging + metric sinks (reference: hivemind logger reuse train_fsdp.py:53,75-76; utils.py:170-204)."""
from __future__ import annotations

import logging
import os
import pickle
from typing import Any, Protocol

_LOGGER: logging.Logger | None = None


def get_logger() -> logging.Logger:
    global _LOGGER
    if _LOGGER is None:
        lg = logging.getLogger("opendiloco_b200")
        if not lg.handlers:
            h = logging.StreamHandler()
            h.setFormatter(logging.Formatter("%(asctime)s [%(levelname)s] %(message)s", "%b %d %H:%M:%S"))
            lg.addHandler(h)
        lg.setLevel(os.environ.get("ODB_LOGLEVEL", "INFO"))
        lg.propagate = False
        _LOGGER = lg
    return _LOGGER


def log_rank(message: str) -> None:
    get_logger().info(f"[rank {os.environ.get('LOCAL_RANK', '0')}] {message}")


class Logger(Protocol):
    def log(self, metrics: dict[str, Any]): ...

    def finish(self): ...


class WandbLogger:
    """wandb sink; resumes the same run id when possible (reference utils.py:179-188)."""

    def __init__(self, project, config, resume: bool = False, **kwargs):
        import wandb

        self._wandb = wandb
        wandb.init(project=project, config=config, resume="auto" if resume else None, **kwargs)

    def log(self, metrics: dict[str, Any]):
        self._wandb.log(metrics)

    def finish(self):
        self._wandb.finish()


class DummyLogger:
    """Collects metric dicts in memory and pickles the list to the path given as ``project`` on finish() — the sink the
    reference's integration tests read back (utils.py:191-204, tests/test_training/test_train.py:70-83)."""

    def __init__(self, project, config, *args, **kwargs):
        self.project, self.config = project, config
        open(project, "a").close()
        self.data: list[dict[str, Any]] = []

    def log(self, metrics: dict[str, Any]):
        self.data.append(metrics)

    def finish(self):
        with open(self.project, "wb") as f:
            pickle.dump(self.data, f)


def make_metric_logger(kind: str, project: str, config: dict, resume: bool = False):
    if kind == "wandb":
        return WandbLogger(project=project, config=config, resume=resume)
    if kind == "dummy":
        return DummyLogger(project=project, config=config)
    raise ValueError(f"unknown metric_logger_type {kind!r} (wandb | dummy)")
