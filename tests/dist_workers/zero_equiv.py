"""Run under torchrun on >= 2 GPUs (NCCL).  The fused ZeRO step kernel (csrc/zero_comm.cu: in-switch reduce-scatter + global
norm + clip + AdamW on the slab + multicast all-gather of the bf16 weights + gradient zeroing, ONE launch) against the
NCCL path (reduce_scatter + norm kernel + all-reduce of the partials + AdamW kernel + all_gather) on identical weights and
gradients, over several steps, for ZeRO-2 and FULL_SHARD."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from opendiloco_b200.models.config import LlamaConfig  # noqa: E402
from opendiloco_b200.models.llama import LlamaForCausalLM  # noqa: E402
from opendiloco_b200.optim.fused import FusedAdamW  # noqa: E402
from opendiloco_b200.parallel import comm  # noqa: E402

comm.init_distributed("nccl")
rank, world = dist.get_rank(), dist.get_world_size()
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
group = dist.group.WORLD
cfg = LlamaConfig(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4, vocab_size=2048)
ok = True


def build(fused: bool, shard_params: bool):
    os.environ["ODB_ZERO_FUSED"] = "1" if fused else "0"
    m = LlamaForCausalLM(cfg, device=dev, seed=11)
    opt = FusedAdamW(m.parameters(), lr=3e-3, betas=(0.9, 0.95), weight_decay=0.1, max_grad_norm=1.0, zero_grad_in_step=True,
                     dp_group=group, shard=True, shard_params=shard_params)
    return m, opt


for shard_params in (False, True):
    ma, oa = build(True, shard_params)
    mb, ob = build(False, shard_params)
    assert (oa._zero is not None) and (ob._zero is None), "fused ZeRO step was not created"
    worst = 0.0
    for step in range(4):
        g = torch.Generator(device=dev).manual_seed(1000 * step + rank)
        grad = torch.randn(ma.arena.numel, device=dev, generator=g) * (0.05 if step != 2 else 5.0)     # step 2 is clipped hard
        ma.arena.grad.copy_(grad)
        mb.arena.grad.copy_(grad)
        oa.step()
        ob.step()
        fa, fb = oa.fv, ob.fv
        lo, hi = fa.lo, fa.hi
        pairs = {"master slab": (fa.flat[lo:hi], fb.flat[lo:hi]), "exp_avg": (oa.exp_avg, ob.exp_avg),
                 "exp_avg_sq": (oa.exp_avg_sq, ob.exp_avg_sq), "stats": (oa.stats, ob.stats),
                 "grads zeroed": (fa.grad, torch.zeros_like(fa.grad))}
        if shard_params:
            pairs["bf16 shard"] = (ma.arena.shadow_shard.float(), mb.arena.shadow_shard.float())
        else:
            pairs["bf16 weights (all ranks' slabs)"] = (ma.arena.shadow.float(), mb.arena.shadow.float())
        for name, (x, y) in pairs.items():
            err = ((x - y).abs().max() / (y.abs().max() + 1e-12)).item()
            worst = max(worst, err)
            # the in-switch reduction and NCCL's ring add the ranks' gradients in different orders: fp32 rounding noise
            # (~1e-7 on the mean gradient) - one bf16 ulp of the largest weight is 4e-3 on this scale
            tol = {"stats": 1e-5, "master slab": 2e-5, "exp_avg": 1e-4, "exp_avg_sq": 1e-4, "grads zeroed": 0.0}.get(name, 8e-3)
            good = err <= tol
            ok &= good
            if not good or (rank == 0 and step == 3):
                print(f"[rank {rank}] shard_params={shard_params} step {step} {name:32s} rel err {err:.2e} {'OK' if good else 'FAIL'}", flush=True)
    if rank == 0:
        print(f"shard_params={shard_params}: worst rel err {worst:.2e}", flush=True)
dist.barrier()
print(f"[rank {rank}] {'ALL OK' if ok else 'SOME FAILED'}", flush=True)
comm.shutdown_distributed()
sys.exit(0 if ok else 1)
