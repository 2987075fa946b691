"""Training-health probes: activation-norm hooks (reference: open_diloco/utils.py:13-67, used at train_fsdp.py:365-387).

``register_metrics_hooks`` attaches forward hooks to every module whose name ends with one of ``target_layers``
(the reference targets ``self_attn`` and ``lm_head``) and accumulates ``||output||_2 / grad_accum`` into
``log_activations["activation/<module name>"]``.  The engine fires these hooks with the o-proj output for
``self_attn`` and - because logits are never materialised - with a norm-only stand-in for ``lm_head`` whose
``.norm()`` comes out of the fused cross-entropy kernel.
"""
from __future__ import annotations

from functools import partial

import torch
from torch.utils.hooks import RemovableHandle

_WRAPPED_NAME_TO_REMOVE = ["_forward_module.", "_fsdp_wrapped_module.", "_orig_mod."]


def _clean_name(name: str) -> str:
    for prefix in _WRAPPED_NAME_TO_REMOVE:
        name = name.replace(prefix, "")
    return name


@torch.no_grad()
def log_activations_hook(_mod, _inp, outp, mod_name: str, gradient_accumulation_steps: int, log_activations: dict) -> None:
    if isinstance(outp, tuple):
        outp = outp[0]
    norm = outp.norm(p=2) / gradient_accumulation_steps
    key = f"activation/{_clean_name(mod_name)}"
    log_activations[key] = norm if key not in log_activations else log_activations[key] + norm


def register_metrics_hooks(model: torch.nn.Module, target_layers: list[str], log_activations: dict,
                           gradient_accumulation_steps: int) -> list[RemovableHandle]:
    handles = []
    for name, mod in model.named_modules():
        if any(name.endswith(layer) for layer in target_layers):
            handles.append(mod.register_forward_hook(partial(log_activations_hook, log_activations=log_activations, mod_name=name,
                                                             gradient_accumulation_steps=gradient_accumulation_steps)))
    return handles


def get_grad_norm(model: torch.nn.Module) -> dict[str, float]:
    """Per-parameter gradient norms (the helper train_diloco_torch.py:24,329 expects from utils)."""
    out = {}
    for name, p in model.named_parameters():
        if p.grad is not None:
            out[f"grad_norm/{_clean_name(name)}"] = p.grad.norm(p=2)
    return out
