"""Base-to-novel inference benchmark (BASELINE config 5): ImageNet-shaped 1000-class text
encoder + image batches of 32 per GPU. Text features are computed once per evaluation — class-sharded over the
ranks (C / world classes each) and all-gathered — then >= 50 image batches run data-parallel through
``MapleEngine.logits`` (the eval path of CustomCLIP.forward, trainers/maple.py:381 / 660-681).

    python tools/infer_bench.py                          # 1 GPU
    torchrun --nproc-per-node 8 tools/infer_bench.py     # 8 GPUs: global batch 256

Prints one JSON line on rank 0: text-tower time (once), images/s over the image batches (CUDA events, max over
ranks), and the same end to end through ``MaPLe.test()`` over a loader of pinned host batches.
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--classes", type=int, default=1000)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--batches", type=int, default=50)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", 1)); rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from types import SimpleNamespace
    from federated_multi_modal_b200 import synth
    from federated_multi_modal_b200.trainers import MaPLe
    cfg = synth.make_cfg(prec="bf16")
    trainer = MaPLe(cfg, client_id=rank, classnames=synth.synthetic_classnames(a.classes))
    eng = trainer.model.engine
    pool = [tuple(t.pin_memory() for t in synth.make_batch(a.batch, a.classes, 50 + 10 * rank + i)) for i in range(4)]
    dpool = [p[0].to(dev) for p in pool]

    class _PinnedLoader:  # the client's test loader: pinned host batches
        def __iter__(self):
            for i in range(a.batches):
                yield {"img": pool[i % 4][0], "label": pool[i % 4][1]}

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn):
        sync()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        sync()
        t = torch.tensor([s.elapsed_time(e)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item() * 1e-3

    shard = world > 1
    eng.logits(dpool[0], cache_text=False, shard_classes=shard)                      # warm-up (buffers, tensor maps)
    t_text_img = timed(lambda: eng.logits(dpool[0], cache_text=False, shard_classes=shard))
    for i in range(3):
        eng.logits(dpool[i % 4])
    t_img = timed(lambda: [eng.logits(dpool[i % 4]) for i in range(a.batches)])
    import contextlib, io
    trainer.dm = SimpleNamespace(test_loader=_PinnedLoader())
    def e2e():  # the call a user makes: MaPLe.test() over the test loader (accuracy read back once)
        with contextlib.redirect_stdout(io.StringIO()):
            trainer.test()
    e2e()
    t_e2e = timed(e2e)
    if rank == 0:
        n = world * a.batch * a.batches
        print(json.dumps({"metric": "MaPLe ViT-B/16 eval images/sec (config 5)", "unit": "images/s",
                          "value": n / t_img, "n_gpus": world, "classes": a.classes, "batch_per_gpu": a.batch,
                          "batches": a.batches, "ms_per_batch": t_img / a.batches * 1e3,
                          "text_tower_plus_one_batch_ms": t_text_img * 1e3,
                          "classes_per_rank_text_tower": a.classes // world if shard else a.classes,
                          "e2e": {"value": n / t_e2e, "unit": "images/s", "ms_per_batch": t_e2e / a.batches * 1e3,
                                  "h2d_bytes_per_batch": pool[0][0].numel() * 4 + a.batch * 8, "d2h_bytes_per_call": 8,
                                  "how": "MaPLe.test() over a loader of pinned host batches"},
                          "dtype": "bf16", "data": "synthetic"}))
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
