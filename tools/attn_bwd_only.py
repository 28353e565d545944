import os, sys
sys.path.insert(0, os.getcwd())
import torch
from federated_multi_modal_b200 import ops
N, T, H = 32, 199, 12
D = H * 64
dev = "cuda"
qkv = torch.randn(N * T, 3 * D, device=dev).to(torch.bfloat16)
out = torch.empty(N * T, D, device=dev, dtype=torch.bfloat16)
do = torch.randn(N * T, D, device=dev).to(torch.bfloat16)
lse = torch.empty(N, H, T, device=dev); delta = torch.empty(N * H * T, device=dev); dqkv = torch.empty_like(qkv)
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
ops.attn_fwd(qkv, out, lse, N, T, H, False)
def timeit(fn):
    for _ in range(3): fn()
    ts = []
    for _ in range(10):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort(); return ts[len(ts) // 2]
cold = timeit(lambda: ops.attn_bwd(qkv, out, do, lse, delta, dqkv, N, T, H, False, impl="fused"))
cold_f = timeit(lambda: ops.attn_fwd(qkv, out, lse, N, T, H, False))
def warm(fn, n=50):
    for _ in range(5): fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) * 1e3 / n
w = warm(lambda: ops.attn_bwd(qkv, out, do, lse, delta, dqkv, N, T, H, False, impl="fused"))
wf = warm(lambda: ops.attn_fwd(qkv, out, lse, N, T, H, False))
print(os.environ.get("MFK_LIB_PATH", "default"), "bwd cold %.1f us, warm back-to-back %.1f us; fwd cold %.1f, warm %.1f" % (cold, w, cold_f, wf))
