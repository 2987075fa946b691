import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opendiloco_b200.ops import attention as A
qkv = torch.randn(32 * 1024, 3 * 1024, device="cuda").to(torch.bfloat16)
for _ in range(3): A.tc_attention_fwd(qkv, 32, 1024, 16, 16)
torch.cuda.synchronize()
