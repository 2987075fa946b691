"""Fused outer step over an NVLink symmetric-memory window (filled in by csrc/outer_comm.cu)."""
from __future__ import annotations


def try_make_fused_outer(opt, compression=None):
    """Return a FusedOuterStep bound to ``opt`` or None when the fused path is unavailable."""
    try:
        from ._fused_outer_impl import FusedOuterStep
    except Exception:
        return None
    return FusedOuterStep.try_create(opt, compression)
