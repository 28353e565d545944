"""Cycle-level phase trace of CTA 0 of the fused attention backward (clock64 stamps)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from federated_multi_modal_b200 import ops, _lib
N, T, H = 32, 199, 12
D = H * 64
dev = "cuda"
qkv = torch.randn(N * T, 3 * D, device=dev).to(torch.bfloat16)
out = torch.empty(N * T, D, device=dev, dtype=torch.bfloat16)
do = torch.randn(N * T, D, device=dev).to(torch.bfloat16)
lse = torch.empty(N, H, T, device=dev); delta = torch.empty(N * H * T, device=dev); dqkv = torch.empty_like(qkv)
ops.attn_fwd(qkv, out, lse, N, T, H, False)
for _ in range(3):
    ops.attn_bwd(qkv, out, do, lse, delta, dqkv, N, T, H, False)
buf = torch.zeros(2 * 64, device=dev, dtype=torch.int64)
_lib.call("mfk_debug_set_attn_trace", buf)
ops.attn_bwd(qkv, out, do, lse, delta, dqkv, N, T, H, False)
torch.cuda.synchronize()
_lib.call("mfk_debug_set_attn_trace", None)
t = buf.cpu().reshape(2, 64)
base = int(t[0, 0])
names_ctl = ["unit start", "loads landed", "S/dP issued"] + sum([[f"blk{k} staged-ready", f"blk{k} mma2(+next SdP) issued"] for k in range(4)], [])
print("control thread (cycles since unit start):")
row = [int(x) - base for x in t[0] if int(x) > 0]
for i, v in enumerate(row[:33]):
    print(f"  {names_ctl[i % 11]:32s} {v:8d}" + ("   <-- next unit" if i % 11 == 0 and i else ""))
print("compute warp 0:")
row = [int(x) - base for x in t[1] if int(x) > 0]
# per block: about to wait for S/dP, S/dP ready, [first half of the arithmetic done, previous block's staged tiles released]
# (not in the very first block), staged
print("  ", row[:60])
names = ["wait", "ready", "math0", "released", "staged"]
r = row[3:]  # first block of the launch has no mma2 wait
for b in range(len(r) // 5):
    x = r[b * 5:(b + 1) * 5]
    print(f"  block {b + 1}: S/dP wait {x[1] - x[0]:5d}  math(1st half) {x[2] - x[1]:5d}  mma2 wait {x[3] - x[2]:5d}  "
          f"stores + 2nd half {x[4] - x[3]:5d}" + (f"  -> next wait {r[(b + 1) * 5] - x[4]:5d}" if (b + 1) * 5 < len(r) else ""))
