"""Micro-benchmark of the LayerNorm kernels on the vision shape (CUDA events, L2 flushed)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from federated_multi_modal_b200 import ops
M, D = int(os.environ.get("LN_M", 6368)), 768
dev = "cuda"
x = torch.randn(M, D, device=dev); g = torch.randn(D, device=dev); b = torch.randn(D, device=dev)
y16 = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
mean = torch.empty(M, device=dev); rstd = torch.empty(M, device=dev)
dy = torch.randn(M, D, device=dev).to(torch.bfloat16)
gin = torch.randn(M, D, device=dev); g16 = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
dg = torch.empty(D, device=dev); db = torch.empty(D, device=dev)
ws = torch.empty(2 * D * ops.ln_bwd_ctas(M), device=dev)
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
def timeit(fn, flush_l2=True):
    for _ in range(3): fn()
    ts = []
    for _ in range(10):
        if flush_l2: flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort(); return ts[len(ts) // 2]
f = lambda: ops.layernorm_fwd(x, g, b, y_bf16=y16, mean=mean, rstd=rstd)
bw = lambda: ops.layernorm_bwd(dy, x, mean, rstd, g, g_in=gin, g_out=gin, g_out_bf16=g16, dgamma=dg, dbeta=db, partial_ws=ws)
bw0 = lambda: ops.layernorm_bwd(dy, x, mean, rstd, g, g_in=gin, g_out=gin, g_out_bf16=g16)
for name, fn, bytes_ in (("ln_fwd", f, M * D * 6), ("ln_bwd(+dgamma,dbeta)", bw, M * D * 20), ("ln_bwd(no param grads)", bw0, M * D * 20)):
    for fl in (True, False):
        t = timeit(fn, fl)
        print(f"{name:24s} L2 {'cold' if fl else 'warm'}: {t:7.1f} us  {bytes_ / t / 1e3:7.1f} GB/s algorithmic")
