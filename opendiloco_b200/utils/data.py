"""Data pipeline: synthetic tokens (fast path), HF streaming text (when a tokenizer/dataset is reachable), collation.

Reference behaviour being reproduced:
  * ``FakeTokenizedDataset``: infinite random token ids in [3, vocab), attention mask all ones (utils.py:155-167),
    used with TEST_VOCAB_SIZE=1024 (train_fsdp.py:66,133-134);
  * causal-LM collation: labels = input_ids with pad positions -> -100 (DataCollatorForLanguageModeling(mlm=False));
  * worker-major sharding: rank = world_rank * world_size + local_rank of galaxy_size * world_size
    (train_fsdp.py:151-156);
  * resumable loader state (torchdata StatefulDataLoader: train_fsdp.py:163-168, ckpt_utils.py:83-87).

``SyntheticTokenLoader`` is the B200-side fast path: batches are produced directly as pinned [B, S] int64 tensors
(by the C++ prefetcher in ``libodbhost.so`` when built) so the H2D copy is a single async DMA per micro-batch.
"""
from __future__ import annotations

from typing import Any, Generator, Iterator

import torch
from torch.utils.data import IterableDataset

TEST_VOCAB_SIZE = 1024
IGNORE_INDEX = -100


class FakeTokenizedDataset(IterableDataset):
    """Infinite random sequences of length ``seq_len`` over ``[3, vocab_size)`` (reference utils.py:155-167).
    ``seed`` (an extension) makes the stream reproducible per rank; None keeps the reference's unseeded behaviour."""

    def __init__(self, seq_len: int, vocab_size: int, seed: int | None = None):
        assert vocab_size > 3, "Vocab size must be greater than 3"
        self.seq_len, self.vocab_size, self.seed = seq_len, vocab_size, seed

    def __iter__(self) -> Generator[dict[str, Any], Any, None]:
        gen = None
        if self.seed is not None:
            gen = torch.Generator().manual_seed(self.seed)
        while True:
            ids = torch.randint(3, self.vocab_size, (self.seq_len,), generator=gen).tolist()
            yield {"input_ids": ids, "attention_mask": [1] * self.seq_len}


def collate_causal_lm(features: list[dict], pad_token_id: int | None = None) -> dict[str, torch.Tensor]:
    """Pad to the longest sequence in the batch and build labels (pad -> -100)."""
    mx = max(len(f["input_ids"]) for f in features)
    pad = pad_token_id if pad_token_id is not None else 0
    ids = torch.full((len(features), mx), pad, dtype=torch.int64)
    mask = torch.zeros((len(features), mx), dtype=torch.int64)
    for i, f in enumerate(features):
        n = len(f["input_ids"])
        ids[i, :n] = torch.as_tensor(f["input_ids"], dtype=torch.int64)
        m = f.get("attention_mask")
        mask[i, :n] = torch.as_tensor(m, dtype=torch.int64) if m is not None else 1
    labels = ids.clone()
    labels[mask == 0] = IGNORE_INDEX
    if pad_token_id is not None:
        labels[ids == pad_token_id] = IGNORE_INDEX
    return {"input_ids": ids, "attention_mask": mask, "labels": labels}


class SyntheticTokenLoader:
    """Resumable iterator of pinned synthetic batches ``{"input_ids", "attention_mask", "labels"}`` of shape [B, S].

    Token law = the reference fake data (uniform over [3, vocab)); the generator is seeded per (seed, rank) and its
    state is part of ``state_dict()`` so checkpoint/resume reproduces the stream exactly."""

    def __init__(self, batch_size: int, seq_len: int, vocab_size: int = TEST_VOCAB_SIZE, seed: int = 0, rank: int = 0,
                 pin_memory: bool | None = None, with_mask: bool = True):
        self.batch_size, self.seq_len, self.vocab_size = batch_size, seq_len, vocab_size
        self.gen = torch.Generator().manual_seed(seed * 1_000_003 + rank)
        self.pin = torch.cuda.is_available() if pin_memory is None else pin_memory
        self.with_mask = with_mask
        self._mask = torch.ones(batch_size, seq_len, dtype=torch.int64)
        self.batches_yielded = 0

    def __iter__(self) -> Iterator[dict[str, torch.Tensor]]:
        return self

    def __next__(self) -> dict[str, torch.Tensor]:
        ids = torch.empty(self.batch_size, self.seq_len, dtype=torch.int64, pin_memory=self.pin)
        torch.randint(3, self.vocab_size, ids.shape, generator=self.gen, out=ids)
        self.batches_yielded += 1
        out = {"input_ids": ids, "labels": ids}
        if self.with_mask:
            out["attention_mask"] = self._mask
        return out

    def state_dict(self) -> dict:
        return {"rng": self.gen.get_state(), "batches_yielded": self.batches_yielded}

    def load_state_dict(self, sd: dict) -> None:
        self.gen.set_state(sd["rng"])
        self.batches_yielded = int(sd["batches_yielded"])


def data_rank(world_rank: int | None, galaxy_size: int | None, world_size: int, rank: int, local_rank: int) -> tuple[int, int]:
    """(shard index, number of shards) with the reference's worker-major rule (train_fsdp.py:151-156)."""
    if galaxy_size is not None and world_rank is not None:
        return world_rank * world_size + local_rank, galaxy_size * world_size
    return rank, world_size


def get_text_dataloader(dataset_name_or_path: str, tokenizer_name: str, seq_length: int, batch_size: int, shard: int,
                        num_shards: int, num_workers: int = 4, pad_to_max: bool = True, c4_tiny: bool = False,
                        seed: int | None = None, split: str = "train"):
    """Streaming text -> tokens -> causal-LM batches (reference get_dataloader, train_fsdp.py:132-168;
    train_diloco_torch.py:201-229).  Needs ``datasets`` + a locally available tokenizer/dataset (no network here)."""
    from datasets import load_dataset
    from datasets.distributed import split_dataset_by_node
    from transformers import AutoTokenizer

    try:
        from torchdata.stateful_dataloader import StatefulDataLoader as Loader
    except Exception:  # pragma: no cover
        from torch.utils.data import DataLoader as Loader

    tok = AutoTokenizer.from_pretrained(tokenizer_name, use_fast=True)
    tok.pad_token = "</s>"
    if c4_tiny:
        ds = load_dataset("PrimeIntellect/c4-tiny", "en")
    else:
        ds = load_dataset(dataset_name_or_path, "en", streaming=True)
    if seed is not None:
        ds = ds.shuffle(seed=seed)

    def tokenize(batch):
        kw = dict(truncation=True, max_length=seq_length)
        if pad_to_max:
            kw["padding"] = "max_length"
        return tok(batch["text"], **kw)

    cols = [c for c in ("text", "timestamp", "url") if c in (ds[split].column_names or ["text", "timestamp", "url"])]
    tokenized = ds.map(tokenize, batched=True, remove_columns=cols)[split]
    sharded = split_dataset_by_node(tokenized, world_size=num_shards, rank=shard)
    pad_id = tok.pad_token_id

    def collate(features):
        return collate_causal_lm(features, pad_id)

    return Loader(sharded, collate_fn=collate, batch_size=batch_size, num_workers=num_workers)


def get_fake_dataloader(seq_length: int, batch_size: int, vocab_size: int = TEST_VOCAB_SIZE, num_workers: int = 0,
                        seed: int | None = None):
    """The reference's --fake-data loader: FakeTokenizedDataset through a (stateful) torch DataLoader."""
    try:
        from torchdata.stateful_dataloader import StatefulDataLoader as Loader
    except Exception:  # pragma: no cover
        from torch.utils.data import DataLoader as Loader
    return Loader(FakeTokenizedDataset(seq_length, vocab_size, seed), collate_fn=collate_causal_lm, batch_size=batch_size,
                  num_workers=num_workers)
