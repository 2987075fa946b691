"""Sharding strategies inside one DiLoCo worker (reference: utils.py:138-152 -> torch FSDP1).

The reference wraps the whole model in ONE FSDP unit (no auto-wrap policy), i.e. one flat parameter whose gradient is
reduce-scattered and whose optimizer state is sharded.  Here that flat parameter is the arena itself:

    NO_SHARD                          flat gradient all-reduce, replicated AdamW                (DDP-equivalent)
    SHARD_GRAD_OP, _HYBRID_SHARD_ZERO2 flat gradient reduce-scatter, AdamW on the rank's slice of (master, m, v,
                                      theta_outer, outer momentum), all-gather of the bf16 compute weights   (ZeRO-2)
    FULL_SHARD, HYBRID_SHARD          same state sharding AND sharded compute weights (ZeRO-3): a rank keeps its slice of the
                                      bf16 weights; the flat parameter is all-gathered before forward, freed after it,
                                      all-gathered again before backward and freed after it - FSDP's behaviour for the
                                      reference's single wrapping unit (``ParamArena.enable_param_sharding``)

"HYBRID" (= shard inside a node, replicate across nodes) coincides with "shard inside a worker, DiLoCo across workers"
in this framework's topology.  Implementation: ``optim.fused.FusedAdamW(dp_group=..., shard=...)``.
"""
from __future__ import annotations

from enum import Enum


class ShardingStrategy(Enum):
    FULL_SHARD = "FULL_SHARD"
    SHARD_GRAD_OP = "SHARD_GRAD_OP"
    NO_SHARD = "NO_SHARD"
    HYBRID_SHARD = "HYBRID_SHARD"
    _HYBRID_SHARD_ZERO2 = "_HYBRID_SHARD_ZERO2"

    @property
    def shards_optimizer_state(self) -> bool:
        return self is not ShardingStrategy.NO_SHARD

    @property
    def shards_parameters(self) -> bool:
        return self in (ShardingStrategy.FULL_SHARD, ShardingStrategy.HYBRID_SHARD)


def get_sharding_strategy(sharding_strategy: str) -> ShardingStrategy:
    try:
        return ShardingStrategy(sharding_strategy)
    except ValueError:
        raise ValueError(f"Invalid sharding_strategy: {sharding_strategy}. Please choose 'FULL_SHARD', 'SHARD_GRAD_OP', "
                         "'NO_SHARD', 'HYBRID_SHARD', or '_HYBRID_SHARD_ZERO2'.") from None
