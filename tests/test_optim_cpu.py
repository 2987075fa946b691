"""FusedAdamW host logic on CPU (reference math path of every kernel): GradScaler on the fused path (K11), clipping
hand-off, hyper-parameter ring.  Reference call sites: train_fsdp.py:390-408, utils.py:124-135."""
import copy

import torch

from opendiloco_b200.optim.fused import FusedAdamW
from opendiloco_b200.utils.training import found_inf_grad


def _models():
    torch.manual_seed(0)
    a = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.Linear(32, 8))
    return a, copy.deepcopy(a)


def test_scaled_step_matches_torch_adamw_and_skips_on_overflow():
    ours, ref = _models()
    opt = FusedAdamW(ours.parameters(), lr=1e-2, betas=(0.9, 0.95), weight_decay=0.1, max_grad_norm=1.0)
    ropt = torch.optim.AdamW(ref.parameters(), lr=1e-2, betas=(0.9, 0.95), weight_decay=0.1)
    scaler = torch.amp.GradScaler("cpu", init_scale=1024.0, growth_interval=1000)
    x = torch.randn(4, 16)
    for step in range(3):
        loss = ours(x).pow(2).mean()
        scaler.scale(loss).backward()                       # gradients carry the 1024x loss scale
        rloss = ref(x).pow(2).mean()
        rloss.backward()
        torch.nn.utils.clip_grad_norm_(ref.parameters(), 1.0)
        ropt.step()
        ropt.zero_grad()
        scaled = next(ours.parameters()).grad.clone()
        opt.unscale_(scaler)                                # nothing is rescaled in memory ...
        assert torch.equal(next(ours.parameters()).grad, scaled)
        opt.step_scaled(scaler)                             # ... 1/scale rides into the fused clip + AdamW
        assert not found_inf_grad(opt, scaler)
        scaler.update()
        opt.zero_grad()
        for p, q in zip(ours.parameters(), ref.parameters()):
            assert torch.allclose(p, q, atol=2e-6), (step, (p - q).abs().max())
    # overflow: the update is skipped, the scaler backs off
    before = [p.detach().clone() for p in ours.parameters()]
    loss = ours(x).pow(2).mean()
    scaler.scale(loss).backward()
    next(ours.parameters()).grad[0, 0] = float("inf")
    opt.unscale_(scaler)
    opt.step_scaled(scaler)
    assert found_inf_grad(opt, scaler)
    scale_before = scaler.get_scale()
    scaler.update()
    assert scaler.get_scale() == scale_before * 0.5
    for p, q in zip(ours.parameters(), before):
        assert torch.equal(p, q)


def test_hp_ring_slots_are_not_reused_while_in_flight():
    ours, _ = _models()
    opt = FusedAdamW(ours.parameters(), lr=1e-3)
    seen = set()
    for i in range(2 * opt.HP_RING):
        for p in ours.parameters():
            p.grad = torch.ones_like(p)
        opt.param_groups[0]["lr"] = 1e-3 * (i + 1)
        slot = opt._hp_slot
        opt.step()
        seen.add(slot)
        assert abs(float(opt._hp[0]) - 1e-3 * (i + 1)) < 1e-9
    assert seen == set(range(opt.HP_RING))
