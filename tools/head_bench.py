"""Training head in isolation: fused single launch vs the six-kernel path (CUDA events, 200 calls each)."""
import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from federated_multi_modal_b200 import ops
dev = "cuda"
for B, C in ((32, 10), (64, 21), (32, 38)):
    E = 512
    g = torch.Generator().manual_seed(0)
    fi, ft = torch.randn(B, E, generator=g).to(dev), torch.randn(C, E, generator=g).to(dev)
    ls = torch.tensor([math.log(1 / 0.07)], device=dev)
    lab = torch.randint(0, C, (B,), generator=g).to(dev)
    logits, loss = torch.empty(B, C, device=dev), torch.empty(1, device=dev)
    dfi, dft = torch.empty(B, E, device=dev), torch.empty(C, E, device=dev)
    ws = torch.empty(ops.head_workspace_floats(B, C, E), device=dev)
    def run(n):
        for _ in range(n):
            ops.head_forward_backward(fi, ft, ls, lab, logits, loss, dfi, dft, ws)
    run(5); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); run(200); e.record(); torch.cuda.synchronize()
    print(f"B={B} C={C}: {s.elapsed_time(e) * 1e3 / 200:.2f} us per training head call")
