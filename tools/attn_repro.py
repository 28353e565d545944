import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from federated_multi_modal_b200 import ops
N, T, H = 2, 199, 12
D = H * 64
qkv = torch.randn(N * T, 3 * D, device="cuda").to(torch.bfloat16)
out = torch.empty(N * T, D, device="cuda", dtype=torch.bfloat16)
do = torch.randn(N * T, D, device="cuda").to(torch.bfloat16)
lse = torch.empty(N, H, T, device="cuda"); delta = torch.empty(N * H * T, device="cuda"); dqkv = torch.empty_like(qkv)
ops.attn_fwd(qkv, out, lse, N, T, H, False)
ops.attn_bwd(qkv, out, do, lse, delta, dqkv, N, T, H, False, impl="fused")
torch.cuda.synchronize()
print("ok", dqkv.float().abs().mean().item())
