"""LayerNorm forward / backward on the vision shape, 12 distinct operand sets (12 x 29 / 78 MB >> L2) launched back to
back like inside a step (PDL-chained), CUDA events around 10 rounds: per-launch time and algorithmic GB/s."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from federated_multi_modal_b200 import ops
M, D, L = int(os.environ.get("LN_M", 6368)), 768, 12
dev = "cuda"
x = torch.randn(L, M, D, device=dev); g = torch.randn(D, device=dev); b = torch.randn(D, device=dev)
y16 = torch.empty(L, M, D, device=dev, dtype=torch.bfloat16)
mean = torch.empty(L, M, device=dev); rstd = torch.empty(L, M, device=dev)
dy = torch.randn(L, M, D, device=dev).to(torch.bfloat16)
gin = torch.randn(L, M, D, device=dev); g16 = torch.empty(L, M, D, device=dev, dtype=torch.bfloat16)
dg = torch.empty(D, device=dev); db = torch.empty(D, device=dev)
ws = torch.empty(L, 2 * D * ops.ln_bwd_ctas(M), device=dev)
def fwd():
    for l in range(L):
        ops.layernorm_fwd(x[l], g, b, y_bf16=y16[l], mean=mean[l], rstd=rstd[l])
def bwd():
    for l in range(L):
        ops.layernorm_bwd(dy[l], x[l], mean[l], rstd[l], g, g_in=gin[l], g_out=gin[l], g_out_bf16=g16[l], dgamma=dg,
                          dbeta=db, partial_ws=ws[l], defer=True)
fwd(); torch.cuda.synchronize()
for name, fn, bytes_ in (("ln_fwd", fwd, M * D * 6), ("ln_bwd", bwd, M * D * 16)):
    fn(); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()          # replayed from a graph: the host's launch rate must not be what is measured
    with torch.cuda.graph(gr):
        fn()
    gr.replay(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        gr.replay()
    e.record(); torch.cuda.synchronize()
    t = s.elapsed_time(e) * 1e3 / (10 * L)
    print(f"{name}: {t:6.2f} us per launch, {bytes_ / t / 1e3:7.1f} GB/s algorithmic ({bytes_ / 1e6:.1f} MB)")
