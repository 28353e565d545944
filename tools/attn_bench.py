"""Micro-benchmark of the attention kernels on the MaPLe vision shape (CUDA events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from federated_multi_modal_b200 import ops
N, T, H = int(os.environ.get("AB_N", 32)), 199, 12
D = H * 64
dev = "cuda"
qkv = torch.randn(N * T, 3 * D, device=dev).to(torch.bfloat16)
out = torch.empty(N * T, D, device=dev, dtype=torch.bfloat16)
do = torch.randn(N * T, D, device=dev).to(torch.bfloat16)
lse = torch.empty(N, H, T, device=dev)
delta = torch.empty(N * H * T, device=dev)
dqkv = torch.empty_like(qkv)
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
def timeit(fn):
    for _ in range(3): fn()
    ts = []
    for _ in range(10):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort(); return ts[len(ts) // 2]
unit = 2.0 * T * T * 64 * N * H  # one T x T x 64 GEMM over all (sequence, head) pairs
for impl in ("mma", "tc", "fused"):
    f = timeit(lambda: ops.attn_fwd(qkv, out, lse, N, T, H, False, impl="tc" if impl == "fused" else impl))
    b = timeit(lambda: ops.attn_bwd(qkv, out, do, lse, delta, dqkv, N, T, H, False, impl=impl))
    print(f"{impl}: fwd {f:7.1f} us ({2 * unit / f / 1e6:6.1f} TFLOP/s algorithmic)   bwd {b:7.1f} us ({5 * unit / b / 1e6:6.1f} TFLOP/s algorithmic)")
