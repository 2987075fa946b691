"""clock64 timeline of CTA (0,0,0) of the tcgen05 attention backward (key tile 0: 8 query tiles, then key tile 7: 1 query tile)."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opendiloco_b200 import _lib
from opendiloco_b200.ops import attention as A
B, S, H = 32, 1024, 16
qkv = torch.randn(B * S, 3 * H * 64, device="cuda").to(torch.bfloat16)
out, lse = A.tc_attention_fwd(qkv, B, S, H, H)
dout = torch.randn_like(out)
for _ in range(2): A.tc_attention_bwd(dout, qkv, out, lse, B, S, H, H)
dbg = torch.zeros(16 * 16, dtype=torch.int64, device="cuda")
lib = _lib.cuda_lib()
lib.odb_attn_bwd_set_dbg.argtypes = [ctypes.c_void_p]
lib.odb_attn_bwd_set_dbg(dbg.data_ptr())
A.tc_attention_bwd(dout, qkv, out, lse, B, S, H, H)
torch.cuda.synchronize()
lib.odb_attn_bwd_set_dbg(None)
d = dbg.cpu().view(16, 16)
t0 = int(d[0, 0])
names = ["ew:loop", "sdp_full", "ld+free", "math", "mma_done", "PdS_out", "dq_ld", "stg_free", "dq_red", "mma:loop", "sdp_issued", "pds_full", "dq_empty"]
print("stamps (cycles since the element-wise loop start of CTA 0), one row per query tile; dq_* columns belong to the previous tile's dQ")
print(" ".join(f"{n:>10s}" for n in names))
for j in range(9):
    print(" ".join(f"{int(d[j, k]) - t0 if int(d[j, k]) else 0:10d}" for k in range(13)))
