"""Fused linear-cross-entropy vs the materialising path at the Llama-150M micro-batch shape (CUDA events, after warm-up).

    python profiles/lce_bench.py [T] [V] [h]
"""
import sys

import torch

from opendiloco_b200.ops import gemm as G
from opendiloco_b200.ops import kernels as K
from opendiloco_b200.ops import tc_gemm as T

Tn, V, h = (int(a) for a in (sys.argv[1:4] + [32768, 32000, 1024][len(sys.argv) - 1:]))
BF = torch.bfloat16
torch.manual_seed(0)
x = torch.randn(Tn, h, device="cuda").to(BF)
w = (torch.randn(V, h, device="cuda") * 0.02).to(BF)
labels = torch.randint(0, V, (Tn,), device="cuda")
gscale = torch.full((1,), 1.0 / Tn, device="cuda")
loss_sum = torch.zeros(1, device="cuda")
planes = T.lce_planes(V)
shift, rowscale = torch.empty(Tn, device="cuda"), torch.empty(Tn, device="cuda")
partials = torch.empty(2 * planes * Tn, device="cuda")
e = torch.empty(Tn, V, device="cuda", dtype=BF)
xs = torch.empty(Tn, h, device="cuda", dtype=BF)
dw = torch.zeros(V, h, device="cuda")
dx = torch.empty(Tn, h, device="cuda", dtype=BF)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()                                  # flush L2 between timed iterations
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return sorted(ts)[len(ts) // 2]


steps = {
    "label_dot": lambda: K.lce_label_dot(x, w, labels, shift),
    "lce_fwd (GEMM + exp epilogue + store)": lambda: T.lce_fwd(x, w, shift, partials, e),
    "lce_fwd eval (no store)": lambda: T.lce_fwd(x, w, shift, partials, None),
    "finalize (+xs, +one-hot scatter)": lambda: K.lce_finalize(partials, planes, labels, gscale, loss_sum, None, rowscale, x, xs, dw),
    "lce_dx (dgrad, MN-major W, scale+gather epilogue)": lambda: T.lce_dx(e, w, rowscale, labels, gscale, dx),
    "wgrad E^T xs": lambda: G.mm_tn_acc(e, xs, dw),
    "old: logits GEMM": lambda: T.linear(x, w, e),
    "old: CE in place": lambda: K.ce_fwd_bwd_(e, labels, gscale, loss_sum),
    "old: dX GEMM (cuBLAS nn)": lambda: torch.mm(e, w, out=dx),
    "plain nn GEMM (ours, MN-major B)": lambda: T.linear_nn(e, w, dx),
}
fl = 2.0 * Tn * V * h
tot_new = tot_old = 0.0
for name, fn in steps.items():
    us = timeit(fn)
    gemm = any(k in name for k in ("GEMM", "lce_fwd", "lce_dx", "wgrad"))
    print(f"{name:52s} {us:9.1f} us" + (f"  {fl / us / 1e6:7.1f} TF/s" if gemm else ""))
new = ["label_dot", "lce_fwd (GEMM + exp epilogue + store)", "finalize (+xs, +one-hot scatter)",
       "lce_dx (dgrad, MN-major W, scale+gather epilogue)", "wgrad E^T xs"]
old = ["old: logits GEMM", "old: CE in place", "plain nn GEMM (ours, MN-major B)", "wgrad E^T xs"]
