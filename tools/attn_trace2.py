"""Debug: extra clock64 stamps of the fused attention backward (build with MFK_DEFS=-DMFK_TRACE2)."""
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
buf = torch.zeros(5 * 64, device=dev, dtype=torch.int64)
_lib.call("mfk_debug_set_attn_trace", buf)
ops.attn_bwd(qkv, out, do, lse, delta, dqkv, N, T, H, False)
torch.cuda.synchronize()
_lib.call("mfk_debug_set_attn_trace", None)
t = buf.cpu().reshape(5, 64)
base = int(t[0, 0])
for i, nm in enumerate(["A (11/unit: start, landed, sdp, 4x(staged-ready, issued))", "compute w0 (wait, ready, [math0, released], staged)",
                        "B (per block: go, committed)", "drain q0 (per key tile: wait, acc_done seen, drained)",
                        "compute w0 unit boundary (last block staged, next unit statistics available)"]):
    print(nm)
    print("   ", [int(x) - base for x in t[i] if int(x) > 0])
