#!/usr/bin/env python
"""Tokenise a text corpus ONCE into shards for ``TokenFileLoader`` (``opendiloco_b200/csrc/host/tokenfile.cc``).

    python scripts/tokenize_corpus.py --input c4/en/c4-train.0000*.json.gz --output-dir /data/c4-tok \
        --tokenizer mistralai/Mistral-7B-v0.1 --shard-tokens 200000000
    torchrun ... -m opendiloco_b200.train_fsdp --dataset-name-or-path "tokens:/data/c4-tok/*.tok" ...

Inputs: ``.txt`` (one document per line), ``.jsonl`` / ``.json`` / ``.json.gz`` with a ``text`` field per line (the C4 layout
that ``scripts/pull-c4.sh`` clones).  Every document is followed by the tokenizer's EOS; documents are concatenated
(the loader cuts fixed windows, no padding tokens are ever trained on).  ``--tokenizer byte`` is a dependency-free
byte-level fallback (ids 3..258, EOS = 2) for smoke tests.
"""
from __future__ import annotations

import argparse
import glob
import gzip
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from opendiloco_b200.utils.data import write_token_shard  # noqa: E402


def documents(path: str):
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rt", encoding="utf-8") as f:
        if ".json" in os.path.basename(path):
            for line in f:
                line = line.strip()
                if line:
                    yield json.loads(line)["text"]
        else:
            for line in f:
                if line.strip():
                    yield line.rstrip("\n")


def make_encoder(name: str):
    if name == "byte":
        return (lambda text: [b + 3 for b in text.encode("utf-8")]), 2
    from transformers import AutoTokenizer

    tok = AutoTokenizer.from_pretrained(name, use_fast=True)
    eos = tok.eos_token_id if tok.eos_token_id is not None else 2
    return (lambda text: tok(text, add_special_tokens=False)["input_ids"]), eos


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--input", nargs="+", required=True, help="files or globs")
    ap.add_argument("--output-dir", required=True)
    ap.add_argument("--tokenizer", default="mistralai/Mistral-7B-v0.1")
    ap.add_argument("--shard-tokens", type=int, default=100_000_000)
    ap.add_argument("--prefix", default="shard")
    a = ap.parse_args()
    files = sorted(f for pat in a.input for f in (glob.glob(pat) or [pat]))
    encode, eos = make_encoder(a.tokenizer)
    os.makedirs(a.output_dir, exist_ok=True)
    buf: list[int] = []
    n_shards = n_tokens = n_docs = 0

    def flush():
        nonlocal buf, n_shards
        if buf:
            write_token_shard(os.path.join(a.output_dir, f"{a.prefix}_{n_shards:05d}.tok"), torch.tensor(buf, dtype=torch.int64))
            n_shards += 1
            buf = []

    for path in files:
        for doc in documents(path):
            ids = encode(doc)
            buf.extend(ids)
            buf.append(eos)
            n_tokens += len(ids) + 1
            n_docs += 1
            if len(buf) >= a.shard_tokens:
                flush()
    flush()
    print(f"{n_docs} documents, {n_tokens} tokens -> {n_shards} shard(s) in {a.output_dir}")


if __name__ == "__main__":
    main()
