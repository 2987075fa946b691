"""Warm the local HF cache with the tokenizer used for real-data runs (reference: scripts/pull-model.py). Needs network."""
from transformers import AutoTokenizer

AutoTokenizer.from_pretrained("mistralai/Mistral-7B-v0.1", use_fast=True)
