"""Federated-round benchmark through the drop-in trainer API (BASELINE.json metric ii, "FedAvg round time (s)"):
``MaPLeFederated.train()`` on the synthetic shapes of BASELINE configs 3 and 4, one process per GPU.

    python tools/fed_round_bench.py --preset c3            # 38 classes, 8 clients, Dirichlet(0.5) label split, B=32
    torchrun --nproc-per-node 4 tools/fed_round_bench.py --preset c4   # 21 classes, 32 clients, B=64, 16 steps/client

c3: a 38 x 160 synthetic image pool is split over the clients by ``dirichlet_label_split(alpha=0.5, seed=0)``;
every client runs floor(n_k / B) steps per round (LOCAL_EPOCHS = 1) — round time = slowest rank.
c4: every client runs exactly 16 steps of batch 64 per round; clients per GPU = 32 / world.
Images are drawn from a small per-rank pool of distinct random tensors (timing does not depend on pixel values;
the labels follow the split). Prints one JSON line on rank 0: median round / local-training / FedAvg seconds over
the rounds after the first (which captures the CUDA graphs), and the aggregate images/s of a round.
"""
import argparse, contextlib, io, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--preset", default="c3", choices=["c3", "c4"])
    ap.add_argument("--clients", type=int, default=0)
    ap.add_argument("--rounds", type=int, default=4)
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--pool", type=int, default=256, help="distinct images per rank")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from federated_multi_modal_b200 import synth
    from federated_multi_modal_b200.fed import clients_of_rank
    from federated_multi_modal_b200.trainers import ClientDataManager, Datum, MaPLeFederated
    from federated_multi_modal_b200.trainers.data_partition import dirichlet_label_split

    if a.preset == "c3":
        C, K, B = 38, a.clients or 8, a.batch or 32
        labels = [c for c in range(C) for _ in range(160)]
        parts = dirichlet_label_split(labels, K, alpha=0.5, seed=0, min_size=B)
        client_labels = [[labels[i] for i in p][: (len(p) // B) * B] for p in parts]
    else:
        C, K, B = 21, a.clients or 32, a.batch or 64
        g = torch.Generator().manual_seed(0)
        client_labels = [torch.randint(0, C, (16 * B,), generator=g).tolist() for _ in range(K)]
    names = synth.synthetic_classnames(C)
    g = torch.Generator().manual_seed(100 + rank)
    pool = [torch.randn(3, 224, 224, generator=g) for _ in range(a.pool)]
    mine = set(clients_of_rank(K, rank, world))
    dms = []
    cfg = synth.make_cfg(prec="bf16")
    cfg.FED.NUM_CLIENTS, cfg.FED.NUM_ROUNDS, cfg.FED.LOCAL_EPOCHS = K, a.rounds, 1
    cfg.DATALOADER = synth._NS(TRAIN_X=synth._NS(BATCH_SIZE=B), TEST=synth._NS(BATCH_SIZE=100))
    cfg.OUTPUT_DIR = ""
    for k in range(K):
        labs = client_labels[k] if k in mine else client_labels[k][:B]  # other ranks' clients: placeholders only
        items = [Datum(impath=f"synthetic://{k}/{i}", label=int(l), classname=names[int(l)], img=pool[(i * 7 + k) % a.pool])
                 for i, l in enumerate(labs)]
        dms.append(ClientDataManager(items, [], [], cfg))
    steps = [len(client_labels[k]) // B for k in range(K)]
    with contextlib.redirect_stdout(io.StringIO()):
        fed = MaPLeFederated(cfg, client_data_managers=dms, classnames=names)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        fed.train()
        torch.cuda.synchronize()
        total = time.perf_counter() - t0
    rt = fed.round_times[1:] or fed.round_times
    med = lambda key: sorted(r[key] for r in rt)[len(rt) // 2]
    mine_t = torch.tensor([med("round_s"), med("local_s"), med("fedavg_s")], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(mine_t, op=dist.ReduceOp.MAX)
    if rank == 0:
        imgs = sum(steps) * B
        print(json.dumps({"metric": "FedAvg round time", "unit": "s", "value": float(mine_t[0]), "preset": a.preset,
                          "n_gpus": world, "clients": K, "clients_per_gpu": K // world, "classes": C, "batch": B,
                          "steps_per_client": {"min": min(steps), "max": max(steps), "sum": sum(steps)},
                          "local_training_s": float(mine_t[1]), "fedavg_s": float(mine_t[2]),
                          "round_images_per_s": imgs / float(mine_t[0]), "rounds_timed": len(rt),
                          "first_round_s": fed.round_times[0]["round_s"], "wall_s": total,
                          "how": "MaPLeFederated.train() (drop-in trainer API, host data loader included), "
                                 "median over rounds after the first, max over ranks"}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
