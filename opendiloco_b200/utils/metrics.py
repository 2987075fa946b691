"""Training-health probes: activation norms and gradient norms (what the reference logs through
open_diloco/utils.py:13-67 and train_fsdp.py:365-387).

The engine does not run ``nn.Module.forward`` per layer, so the probes are organised around one accumulator object,
``ActivationNormProbe``: it owns the metric dict, knows which module names it watches, strips wrapper prefixes from names once
at attach time and exposes the callable the engine fires (``LlamaEngine._fire_hooks``) with the o-proj output of every
``self_attn`` and - because logits are never materialised - a norm-only stand-in for ``lm_head`` whose ``.norm()`` comes out
of the fused linear-cross-entropy epilogue.  ``register_metrics_hooks`` / ``log_activations_hook`` keep the reference's
call signatures on top of it.
"""
from __future__ import annotations

import torch
from torch.utils.hooks import RemovableHandle

WRAPPER_PREFIXES = ("_forward_module.", "_fsdp_wrapped_module.", "_orig_mod.")     # FSDP / torch.compile name decorations


def plain_name(name: str) -> str:
    for prefix in WRAPPER_PREFIXES:
        name = name.replace(prefix, "")
    return name


class ActivationNormProbe:
    """Accumulates ``||output||_2 / gradient_accumulation_steps`` per watched module into ``sink["activation/<name>"]``."""

    def __init__(self, sink: dict, gradient_accumulation_steps: int):
        self.sink, self.scale = sink, 1.0 / float(gradient_accumulation_steps)

    @torch.no_grad()
    def record(self, mod_name: str, output) -> None:
        out = output[0] if isinstance(output, tuple) else output
        key = "activation/" + plain_name(mod_name)
        value = out.norm(p=2) * self.scale
        prev = self.sink.get(key)
        self.sink[key] = value if prev is None else prev + value

    def hook_for(self, mod_name: str):
        def hook(_module, _inputs, output):
            self.record(mod_name, output)

        return hook

    def attach(self, model: torch.nn.Module, suffixes) -> list[RemovableHandle]:
        suffixes = tuple(suffixes)
        return [mod.register_forward_hook(self.hook_for(name)) for name, mod in model.named_modules() if name.endswith(suffixes)]


def log_activations_hook(_mod, _inp, outp, mod_name: str, gradient_accumulation_steps: int, log_activations: dict) -> None:
    """Reference-signature forward hook (utils.py:25-41): one-shot use of the probe."""
    ActivationNormProbe(log_activations, gradient_accumulation_steps).record(mod_name, outp)


def register_metrics_hooks(model: torch.nn.Module, target_layers: list[str], log_activations: dict,
                           gradient_accumulation_steps: int) -> list[RemovableHandle]:
    """Watch every module whose name ends with one of ``target_layers`` (the reference watches ``self_attn`` and ``lm_head``,
    train_fsdp.py:65,365-372); returns the handles to remove afterwards."""
    return ActivationNormProbe(log_activations, gradient_accumulation_steps).attach(model, target_layers)


def get_grad_norm(model: torch.nn.Module) -> dict[str, float]:
    """Per-parameter gradient norms (the helper train_diloco_torch.py:24,329 expects from utils)."""
    return {"grad_norm/" + plain_name(name): p.grad.norm(p=2) for name, p in model.named_parameters() if p.grad is not None}
