import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opendiloco_b200.ops import attention as A
B, S, H = 32, 1024, 16
qkv = torch.randn(B * S, 3 * H * 64, device="cuda").to(torch.bfloat16)
out, lse = A.tc_attention_fwd(qkv, B, S, H, H)
dout = torch.randn_like(out)
for _ in range(2): A.tc_attention_bwd(dout, qkv, out, lse, B, S, H, H)
torch.cuda.synchronize()
