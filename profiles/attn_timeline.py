import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opendiloco_b200.ops import attention as A
qkv = torch.randn(32 * 1024, 3 * 1024, device="cuda").to(torch.bfloat16)
for _ in range(2): A.tc_attention_fwd(qkv, 32, 1024, 16, 16)
dbg = torch.zeros(8 * 16, dtype=torch.int64, device="cuda")
A.tc_attention_fwd(qkv, 32, 1024, 16, 16, dbg=dbg)
torch.cuda.synchronize()
d = dbg.cpu().view(8, 16)
t0 = int(d[0, 0])
names = ["sm:loop", "sm:s_full", "sm:ld_done", "sm:maxbar", "sm:exp_done", "sm:o_done+resc", "sm:P+fence", "sm:arrived", "mma:k_full", "mma:qk_issued", "mma:p_full", "mma:pv_issued"]
print("stamps (cycles since softmax loop start of CTA 0), one row per KV tile")
print(" ".join(f"{n:>14s}" for n in names))
for j in range(8):
    print(" ".join(f"{int(d[j, k]) - t0:14d}" for k in range(12)))
