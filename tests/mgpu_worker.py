"""Worker for multi-rank tests (launched by torchrun, nccl on GPUs or gloo on CPU).
Checks FedAvgExchange: gather order, validity handling, and (CUDA) bit-exact reduction vs the oracle."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from federated_multi_modal_b200.fed import FedAvgExchange, clients_of_rank  # noqa: E402
from oracle.maple_cpu import fedavg_oracle  # noqa: E402


def client_tensor(k, n):
    g = torch.Generator().manual_seed(1000 + k)
    return torch.randn(n, generator=g)


def main():
    backend = sys.argv[1]
    transport = sys.argv[2] if len(sys.argv) > 2 else "auto"
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    if backend == "nccl":
        torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
        dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
        dist.init_process_group("nccl", device_id=dev)
    else:
        dev = torch.device("cpu")
        dist.init_process_group("gloo")
    n, k_local = 4096 * 3 + 64, 2
    K = k_local * world
    mine = clients_of_rank(K, rank, world)
    assert mine == list(range(rank * k_local, (rank + 1) * k_local))
    ex = FedAvgExchange(n, k_local, dev, transport=transport)
    for rnd in range(2):
        for j, k in enumerate(mine):
            t = client_tensor(k + 100 * rnd, n)
            ok = True
            if rnd == 1 and k == 1:
                t[5] = float("nan")          # invalid weights -> must be excluded everywhere
            if rnd == 1 and k == K - 1:
                ok = False                    # failed local training
            ex.publish(j, t.to(dev), ok=ok, n_samples=10 + k)
        rows = ex.gather()
        assert len(rows) == K
        for k in range(K):                    # rank-major gather order == client order
            want = client_tensor(k + 100 * rnd, n)
            got = rows[k].cpu()
            if rnd == 1 and k == 1:
                assert torch.isnan(got[5])
                got = got.clone(); got[5] = 0.0; want[5] = 0.0
            assert torch.equal(got, want), (rank, rnd, k)
        status = ex.status.cpu()
        assert [float(s) for s in status[:, 1]] == [10.0 + k for k in range(K)]
        if backend == "nccl":
            for weighted in (False, True):
                m32, m16, valid, bad = ex.reduce(rows, weighted=weighted)
                expect_valid = [k for k in range(K) if not (rnd == 1 and k in (1, K - 1))]
                assert valid == expect_valid, (valid, expect_valid)
                ref_rows = [client_tensor(k + 100 * rnd, n) for k in expect_valid]
                w = [10.0 + k for k in expect_valid] if weighted else None
                r32, r16 = fedavg_oracle(ref_rows, w)
                assert torch.equal(m32.cpu(), r32), (rank, rnd, weighted)
                assert torch.equal(m16.cpu(), r16)
                # identical on every rank
                chk = m32.double().sum().reshape(1).clone()
                lst = [torch.zeros_like(chk) for _ in range(world)]
                dist.all_gather(lst, chk)
                assert all(torch.equal(x, lst[0]) for x in lst)
        else:
            ex.status.cpu()
        dist.barrier()
    if backend == "nccl":
        # config 5: classes sharded over ranks for the text tower, features all-gathered once
        from helpers import customclip_state_dict
        from federated_multi_modal_b200 import synth
        from federated_multi_modal_b200.engine import MapleEngine
        C = 16 * world
        sd, tok = customclip_state_dict(C, layers=2)
        eng = MapleEngine(sd, tok, device=str(dev))
        img, _ = synth.make_batch(2, C, 11 + rank)
        a = eng.logits(img.to(dev), cache_text=False, shard_classes=True)
        b = eng.logits(img.to(dev), cache_text=False, shard_classes=False)
        assert torch.equal(a, b), (a - b).abs().max()
    if backend == "nccl":
        # end-to-end: MaPLeFederated.train() with 2 clients per rank, one round; every rank must end with the same
        # global arena, equal to the fixed-order mean of ALL clients' published arenas
        import contextlib, io
        from federated_multi_modal_b200.trainers import ClientDataManager, MaPLeFederated
        from federated_multi_modal_b200.trainers.client_datamanager import synthetic_client_items
        Cn, Kc = 4, 2 * world
        cfg = synth.make_cfg()
        cfg.FED.NUM_CLIENTS, cfg.FED.NUM_ROUNDS, cfg.FED.LOCAL_EPOCHS = Kc, 1, 1
        cfg.DATALOADER = synth._NS(TRAIN_X=synth._NS(BATCH_SIZE=2), TEST=synth._NS(BATCH_SIZE=4))
        cfg.OUTPUT_DIR = ""
        names = synth.synthetic_classnames(Cn)
        pool = synthetic_client_items(Cn, Kc, seed=3, classnames=names)          # Cn*Kc images
        dms = [ClientDataManager(pool[k::Kc][:2] + pool[k::Kc][2:4], [], [], cfg) for k in range(Kc)]
        with contextlib.redirect_stdout(io.StringIO()):
            fed = MaPLeFederated(cfg, client_data_managers=dms, classnames=names)
            assert [t.client_id for t in fed.clients] == clients_of_rank(Kc, rank, world)
            captured = {}
            orig = fed._aggregate
            def spy():
                captured["rows"] = [r.clone().cpu() for r in fed.exchange.gather()]
                return orig()
            fed._aggregate = spy
            fed.train()
        assert len(captured["rows"]) == Kc
        m32, m16 = fedavg_oracle(captured["rows"])
        assert torch.equal(fed.last_mean_fp32.cpu(), m32)
        assert torch.equal(fed.global_arena.cpu(), m16.float())
        chk = fed.global_arena.double().sum().reshape(1).clone()
        lst = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(lst, chk)
        assert all(torch.equal(x, lst[0]) for x in lst)
        # local training really differs between clients (different data) before it is averaged
        assert not torch.equal(captured["rows"][0], captured["rows"][Kc - 1])
    if rank == 0:
        print(f"MGPU_OK backend={backend} world={world} transport={ex.transport}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
