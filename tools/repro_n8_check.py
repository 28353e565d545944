"""Single-GPU replay of what each rank of `bench.py --gpus 8` trains (client id / batch seeds of ranks 0..7), followed by
the bench's FedAvg check on the 8 resulting arenas: are all parameters finite, and does the reduce kernel equal the
oracle on them?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from federated_multi_modal_b200 import ops
from oracle.maple_cpu import fedavg_oracle

dev = torch.device("cuda", 0); torch.cuda.set_device(0)
steps = int(os.environ.get("STEPS", 50))
rows = []
for rank in range(8):
    os.environ["RANK"] = str(rank)
    t = bench.make_trainer(dev, graph=True)
    eng = t.model.engine
    pool = bench.host_batches(4, bench.B_PER_GPU, 1000 * rank)
    dev_pool = [(b["img"].to(dev), b["label"].to(dev)) for b in pool]
    for i in range(steps):
        t.step_async(*dev_pool[i % 4])
    loss, norm, flag = t.read_step_result()
    p = eng.params[: eng.n_update]
    nbad = int((~torch.isfinite(p)).sum())
    print(f"rank {rank}: loss {loss:.5f} grad-norm {norm:.4f} flag {flag} non-finite params {nbad} "
          f"max|p| {p.abs().max().item():.4f}", flush=True)
    mine = p.clone()
    mine.mul_(1.0 + 0.03125 * rank).add_(1e-3 * rank)
    rows.append(mine)
    del t, eng
for weighted in (False, True):
    K, n = 8, rows[0].numel()
    ptrs = torch.tensor([r.data_ptr() for r in rows], dtype=torch.int64, device=dev)
    w = torch.tensor([10.0 + k for k in range(K)], dtype=torch.float32, device=dev) if weighted else None
    div = float(sum(10.0 + k for k in range(K))) if weighted else float(K)
    o32 = torch.empty(n, device=dev); o16 = torch.empty(n, device=dev, dtype=torch.float16)
    flags = torch.zeros(K, device=dev, dtype=torch.int32)
    ops.fedavg_reduce(ptrs, w, div, K, n, False, o32, o16, flags)
    r32, r16 = fedavg_oracle([r.cpu() for r in rows], [10.0 + k for k in range(K)] if weighted else None)
    d = (o32.cpu() != r32)
    print(f"weighted={weighted}: fp32 equal {torch.equal(o32.cpu(), r32)} fp16 equal {torch.equal(o16.cpu(), r16)} "
          f"flags {flags.tolist()} mismatches {int(d.sum())}", flush=True)
    if d.any():
        i = int(d.nonzero()[0])
        print("  first mismatch at", i, o32[i].item(), r32[i].item(), [r[i].item() for r in rows])
