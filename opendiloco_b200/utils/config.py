"""CLI -> nested config, with the flag grammar the reference's launch commands use:

    --key value | --key=value      dashes and underscores interchangeable (--per-device-train-batch-size / --per_device_...)
    --a.b value                    dotted nesting (--hv.local-steps 500, --ckpt.interval 8000)
    --flag                         bare flag => True          --no-flag  => False   (--no-torch-compile)
    88_000                         python-style int literals are accepted by pydantic

``BaseConfig`` is a strict pydantic model (unknown flags are errors, like pydantic_config); validation/coercion of
values (ints, floats, Literals, enums, ``str | bool | None``) is pydantic's.
"""
from __future__ import annotations

import re
import sys
from typing import Any

from pydantic import BaseModel, ConfigDict


class BaseConfig(BaseModel):
    model_config = ConfigDict(extra="forbid", validate_default=False, arbitrary_types_allowed=True)


def _norm(key: str) -> str:
    return key.replace("-", "_")


_INT_WITH_UNDERSCORES = re.compile(r"^[+-]?\d+(_\d+)+$")


def _set(d: dict, dotted: str, value: Any) -> None:
    if isinstance(value, str) and _INT_WITH_UNDERSCORES.match(value):
        value = value.replace("_", "")          # 88_000 (R/README.md:122)
    if isinstance(value, str) and value.lower() in ("true", "false"):
        value = value.lower() == "true"
    parts = [_norm(p) for p in dotted.split(".")]
    for p in parts[:-1]:
        d = d.setdefault(p, {})
        if not isinstance(d, dict):
            raise ValueError(f"flag --{dotted} conflicts with a scalar flag of the same prefix")
    d[parts[-1]] = value


def parse_argv(argv: list[str] | None = None) -> dict:
    """Parse ``sys.argv[1:]`` (or ``argv``) into a nested dict of raw strings / bools."""
    args = list(sys.argv[1:] if argv is None else argv)
    out: dict = {}
    i = 0
    while i < len(args):
        a = args[i]
        if not a.startswith("--"):
            raise ValueError(f"expected a --flag, got {a!r}")
        body = a[2:]
        if "=" in body:
            k, v = body.split("=", 1)
            _set(out, k, v)
            i += 1
            continue
        nxt = args[i + 1] if i + 1 < len(args) else None
        is_value = nxt is not None and not (nxt.startswith("--") and not _looks_numeric(nxt))
        if is_value:
            _set(out, body, nxt)
            i += 2
        else:
            head, _, leaf = body.rpartition(".")
            leaf_n = _norm(leaf)
            if leaf_n.startswith("no_"):
                _set(out, (head + "." if head else "") + leaf_n[3:], False)
            else:
                _set(out, body, True)
            i += 1
    return out


def _looks_numeric(s: str) -> bool:
    try:
        float(s)
        return True
    except ValueError:
        return False
