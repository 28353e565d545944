"""Round-end FedAvg exchange alone at world size N: transports p2p / p2p_sharded / nccl, 1 and 4 clients per GPU,
the real trainable-arena size (13 856 768 fp32). CUDA events, max over ranks; bit-exact check against the oracle.

    torchrun --nproc-per-node 8 tools/exchange_bench.py
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from federated_multi_modal_b200.fed import FedAvgExchange
from oracle.maple_cpu import fedavg_oracle  # checker only

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 13856768
out = {}
for kl in (1, 4):
    g = torch.Generator().manual_seed(7)
    base = torch.randn(n, generator=g)
    mine = [(base * (1.0 + 0.01 * (rank * kl + j))).to(dev) for j in range(kl)]
    for tr in ("p2p", "p2p_sharded", "nccl"):
        ex = FedAvgExchange(n, kl, dev, transport=tr, strict_transport=True)
        def once():
            for j in range(kl):
                ex.publish(j, mine[j])
            return ex.reduce(ex.gather())
        m32 = once()[0]
        K = kl * world
        ref = fedavg_oracle([base * (1.0 + 0.01 * k) for k in range(K)])[0] if rank == 0 else None
        ok = torch.equal(m32.cpu(), ref) if rank == 0 else True
        dist.barrier(); torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(10):
            once()
        e.record()
        dist.barrier(); torch.cuda.synchronize()
        t = torch.tensor([s.elapsed_time(e) / 10], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[f"{tr}_k{kl}"] = {"ms": t.item(), "bitexact_rank0": ok,
                              "nvlink_rx_gb_per_rank": ((world - 1) * kl * n * 4 if tr == "p2p" else
                                                        (world - 1) / world * n * (kl * 4 if tr == "p2p_sharded" else kl * 4 * world / (world)) ) / 1e9}
        del ex
if rank == 0:
    print(json.dumps({"world": world, "elements": n, "exchange": out}))
dist.barrier()
dist.destroy_process_group()
