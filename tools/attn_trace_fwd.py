"""Cycle-level phase trace of CTA 0 / softmax warp 0 of the tcgen05 attention forward (clock64 stamps)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from federated_multi_modal_b200 import ops, _lib
N, T, H = 32, 199, 12
D = H * 64
qkv = torch.randn(N * T, 3 * D, device="cuda").to(torch.bfloat16)
out = torch.empty(N * T, D, device="cuda", dtype=torch.bfloat16)
lse = torch.empty(N, H, T, device="cuda")
for _ in range(3):
    ops.attn_fwd(qkv, out, lse, N, T, H, False)
buf = torch.zeros(2 * 64, device="cuda", dtype=torch.int64)
_lib.call("mfk_debug_set_attn_trace", buf)
ops.attn_fwd(qkv, out, lse, N, T, H, False)
torch.cuda.synchronize()
_lib.call("mfk_debug_set_attn_trace", None)
t = [int(x) for x in buf.cpu()[:60] if int(x) > 0]
# stamp order per loop iteration g: wait S(g), S ready(g), pass1(g), pass2(g) [, PV(g-1) ready, epilogue(g-1) done]
names = ["wait S(g)", "S(g) ready", "pass1(g) max", "pass2(g) exp+stage", "PV(g-1) ready", "epilogue(g-1) done"]
base = t[0]
idx = [(0, k) for k in range(4)] + [(g, k) for g in range(1, 12) for k in range(6)]
for i, v in enumerate(t):
    g, k = idx[i]
    if k == 0:
        print(f"--- g = {g}")
    print(f"  {names[k]:20s} {v - base:8d}  (+{v - (t[i - 1] if i else base)})")
