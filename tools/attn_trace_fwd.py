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
# stamp order of softmax group 0 per unit: wait S, S ready, pass1 (row max), pass2 (exp, P staged), PV ready, epilogue done
names = ["wait S", "S ready", "pass1 max", "pass2 exp+stage", "PV ready", "epilogue done"]
base = t[0]
for i, v in enumerate(t):
    if i % 6 == 0:
        print(f"--- unit {i // 6} (group 0 = query tile 0)")
    print(f"  {names[i % 6]:18s} {v - base:8d}  (+{v - (t[i - 1] if i else base)})")
