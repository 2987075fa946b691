"""Small training utilities with the reference's semantics (reference: open_diloco/utils.py:124-152, train_fsdp.py:255-260)."""
from __future__ import annotations

import hashlib
import math

import torch


def found_inf_grad(optimizer: torch.optim.Optimizer, scaler) -> bool:
    """True if ``scaler`` recorded a non-finite gradient for ``optimizer`` during the last unscale/step
    (reads GradScaler's per-optimizer bookkeeping, like the reference utils.py:124-135)."""
    if scaler is None or not scaler.is_enabled():
        return False
    state = scaler._per_optimizer_states.get(id(optimizer))
    if not state or not state.get("found_inf_per_device"):
        return False
    return sum(float(v.item()) for v in state["found_inf_per_device"].values()) > 0


def cosine_schedule_with_warmup_lambda(num_warmup_steps: int, num_training_steps: int, num_cycles: float = 0.5):
    """lambda(step) of HF get_cosine_schedule_with_warmup (ENV/transformers/optimization.py:134-140)."""

    def fn(step: int) -> float:
        if step < num_warmup_steps:
            return float(step) / float(max(1, num_warmup_steps))
        progress = float(step - num_warmup_steps) / float(max(1, num_training_steps - num_warmup_steps))
        return max(0.0, 0.5 * (1.0 + math.cos(math.pi * float(num_cycles) * 2.0 * progress)))

    return fn


def get_cosine_schedule_with_warmup(optimizer, num_warmup_steps: int, num_training_steps: int, last_epoch: int = -1):
    """Linear warm-up then cosine decay to zero, stepped once per optimizer step on the INNER optimizer
    (reference: train_fsdp.py:255-260,407 ; train_diloco_torch.py:189-193,327)."""
    return torch.optim.lr_scheduler.LambdaLR(
        optimizer, cosine_schedule_with_warmup_lambda(num_warmup_steps, num_training_steps), last_epoch)


def hash_tensor_content(a: torch.Tensor, max_size: int = 1000) -> str:
    """Debug fingerprint: md5 of the rounded top-left sqrt(max_size) block (reference utils.py:70-80)."""
    b = int(max_size ** 0.5)
    block = a[:b, :b].flatten() if a.dim() >= 2 else a.flatten()[:max_size]
    txt = ",".join(f"{float(x):.4f}" for x in block[:max_size])
    return hashlib.md5(txt.encode("utf-8")).hexdigest()
