"""GPU: the drop-in trainer surface (CustomCLIP autograd bridge, MaPLe.forward_backward, MaPLeFederated)."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu

from helpers import load_golden
from federated_multi_modal_b200 import synth

if torch.cuda.is_available():
    from federated_multi_modal_b200.trainers import (ClientDataManager, CustomCLIP, MaPLe, MaPLeFederated)
    from federated_multi_modal_b200.trainers.client_datamanager import synthetic_client_items
    from federated_multi_modal_b200.trainers.data_partition import dirichlet_label_split
    from federated_multi_modal_b200.clip import build_model
    from oracle.maple_cpu import fedavg_oracle

DD = {"trainer": "MaPLe", "vision_depth": 0, "language_depth": 0, "vision_ctx": 0, "language_ctx": 0,
      "maple_length": 2}


def _custom_clip(C=10):
    clip = build_model(synth.random_clip_state_dict(0), DD)
    m = CustomCLIP(synth.make_cfg(), synth.synthetic_classnames(C), clip)
    m.prompt_learner.load_state_dict(synth.random_prompt_learner_state(1), strict=False)
    for p in m.parameters():
        p.requires_grad_(False)
    for n, p in m.named_parameters():
        if "prompt_learner" in n or "transformer.resblocks.11" in n:
            p.requires_grad_(True)
    for _, mod in m.named_modules():
        if isinstance(mod, torch.nn.LayerNorm):
            for p in mod.parameters():
                p.requires_grad_(True)
    return m.cuda()


def test_customclip_autograd_bridge_and_eval():
    G = load_golden("c1_fp32.pt")
    img, lab = synth.make_batch(4, 10, 123)
    m = _custom_clip()
    m.eval()
    lg = m(img.cuda())
    # the module holds the reference's fp16 prompt-learner dtypes, so compare at the bf16 tolerance
    assert (lg.cpu() - G["logits_eval"]).abs().max().item() < 2e-2 * G["logits_eval"].abs().max().item()
    m.train()
    loss = m(img.cuda(), lab.cuda())
    assert loss.requires_grad and abs(loss.item() - G["loss"].item()) < 2e-2 * G["loss"].item()
    loss.backward()
    got = {n: p.grad for n, p in m.named_parameters() if p.grad is not None}
    assert len(got) == 145
    for n, g in got.items():
        assert torch.equal(g.float(), m.engine.g[n].to(g.dtype).float()), n
    # a torch optimiser step changes the parameters; the engine picks the new values up
    opt = torch.optim.SGD([p for p in m.parameters() if p.requires_grad], lr=0.01)
    opt.step()
    l2 = m(img.cuda(), lab.cuda())
    assert l2.item() != loss.item()


def _trainer(graph, B=4, C=10):
    cfg = synth.make_cfg()
    cfg.USE_CUDA_GRAPH = graph
    t = MaPLe(cfg, client_id=0, classnames=synth.synthetic_classnames(C))
    t.model.prompt_learner.load_state_dict(synth.random_prompt_learner_state(1), strict=False)
    t.model.load_state_dict(torch.nn.Module.state_dict(t.model))  # push the loaded values into the engine
    return t


def test_maple_forward_backward_graph_equals_eager_and_learns():
    img, lab = synth.make_batch(4, 10, 123)
    batch = {"img": img.pin_memory(), "label": lab.pin_memory()}
    losses = {}
    params = {}
    for graph in (False, True):
        t = _trainer(graph)
        t.model.train()
        losses[graph] = [t.forward_backward(batch)["loss"] for _ in range(4)]
        params[graph] = t.model.engine.params.clone()
        assert len(t.grad_norms) == 4 and all(0 < g <= 1.0 for g in t.grad_norms)
    assert losses[True] == losses[False]            # CUDA-graph replay is bit-identical to eager launches
    assert torch.equal(params[True], params[False])
    assert losses[True][-1] < losses[True][0]       # the step trains
    # state_dict() reflects the fused optimiser's updates, in the reference's dtypes
    sd = t.model.state_dict()
    assert sd["prompt_learner.ctx"].dtype == torch.float16
    assert torch.equal(sd["prompt_learner.compound_prompts_text_parameters.0"],
                       t.model.engine.p["prompt_learner.compound_prompts_text_parameters.0"])


def test_training_is_bit_reproducible_at_bench_batch():
    """Batch 32 (M = 6368 rows: split-K GEMM tails, multi-unit attention CTAs): two independent trainers fed the
    same batches end with bit-identical gradients and parameters — no atomics / races / uninitialised reads."""
    batches = [synth.make_batch(32, 10, 7 + i) for i in range(2)]
    outs = []
    for rep in range(2):
        t = _trainer(False)
        t.model.train()
        for i in range(3):
            img, lab = batches[i % 2]
            t.forward_backward({"img": img.pin_memory(), "label": lab.pin_memory()})
        torch.cuda.synchronize()
        outs.append((t.model.engine.grads.clone(), t.model.engine.params.clone()))
    assert torch.equal(outs[0][0], outs[1][0])
    assert torch.equal(outs[0][1], outs[1][1])


def test_gpu_augment_loader_feeds_trainer():
    """ClientDataManager with the GpuAugment transform: raw uint8 items (EuroSAT-sized 64x64) -> random_resized_crop
    + flip + normalize on the device inside parse_batch_train -> one fused training step."""
    from federated_multi_modal_b200.trainers import Datum, GpuAugment
    g = torch.Generator().manual_seed(0)
    names = synth.synthetic_classnames(10)
    items = [Datum(impath=f"synthetic://{i}", label=i % 10, classname=names[i % 10],
                   img=torch.randint(0, 256, (3, 64, 64), generator=g, dtype=torch.uint8)) for i in range(16)]
    cfg = synth.make_cfg()
    dm = ClientDataManager(items, items[:4], items[:4], cfg, custom_tfm_train=GpuAugment(seed=1))
    batch = next(iter(dm.train_loader))
    assert batch["img_u8"].dtype == torch.uint8 and batch["rrc_box"].shape == (batch["img_u8"].shape[0], 4)
    t = _trainer(False)
    t.model.train()
    x, y, _ = t.parse_batch_train(batch)
    assert x.shape == (batch["img_u8"].shape[0], 3, 224, 224) and x.dtype == torch.float32 and x.is_cuda
    assert torch.isfinite(x).all() and x.abs().max().item() < 3.0   # CLIP-normalised pixel range
    out = t.forward_backward(batch)
    assert out["loss"] > 0


def test_maple_rejects_bad_input_like_reference():
    t = _trainer(False)
    img, lab = synth.make_batch(4, 10, 123)
    img[0, 0, 0, 0] = float("nan")
    with pytest.raises(ValueError, match="NaN"):
        t.forward_backward({"img": img, "label": lab})


def test_federated_rounds_single_gpu_bit_exact_average():
    C, K = 10, 2
    cfg = synth.make_cfg()
    cfg.FED.NUM_CLIENTS, cfg.FED.NUM_ROUNDS, cfg.FED.LOCAL_EPOCHS = K, 2, 1
    cfg.DATALOADER = synth._NS(TRAIN_X=synth._NS(BATCH_SIZE=4), TEST=synth._NS(BATCH_SIZE=8))
    cfg.OUTPUT_DIR = ""
    names = synth.synthetic_classnames(C)
    pool = synthetic_client_items(C, 2, seed=0, classnames=names)  # 20 images
    parts = dirichlet_label_split([it.label for it in pool], K, alpha=0.5, seed=0, min_size=4)
    test_items = pool[:8]
    dms = [ClientDataManager([pool[i] for i in p], [], test_items, cfg) for p in parts]
    fed = MaPLeFederated(cfg, client_data_managers=dms, classnames=names)
    assert len(fed.clients) == K
    # co-located clients share the frozen weights and workspaces
    e0, e1 = fed.clients[0].model.engine, fed.clients[1].model.engine
    assert e0.vis.w[0]["mlp.c_fc.w"].data_ptr() == e1.vis.w[0]["mlp.c_fc.w"].data_ptr()
    assert e0.params.data_ptr() != e1.params.data_ptr()
    # instrument one aggregation: capture what the clients publish
    captured = {}
    orig = fed._aggregate
    def spy():
        captured["rows"] = [r.clone().cpu() for r in fed.exchange.gather()]
        return orig()
    fed._aggregate = spy
    fed.train()
    assert fed.nan_stats["total_updates"] == 2 and len(fed.round_times) == 2
    n = e0.n_update
    mean32, mean16 = fedavg_oracle(captured["rows"])
    assert torch.equal(fed.last_mean_fp32.cpu(), mean32)            # FedAvg'd tensors bit-exact in fp32
    assert torch.equal(fed.global_arena.cpu(), mean16.float())      # and the reference's .half() applied
    for t in fed.clients:                                           # broadcast reached every client
        assert torch.equal(t.model.engine.params[:n], fed.global_arena)
        assert not t.model.engine.mom_initialized                   # optimiser state dropped
    # reference-signature utilities on full state_dicts
    sds = [{k: v.clone() for k, v in t.model.state_dict().items()} for t in fed.clients]
    assert fed.check_weights_valid(sds[0])
    avg = fed.safe_average_weights(sds, K)
    assert set(avg) == set(sds[0]) and all(v.dtype == torch.float16 for v in avg.values())
    k = "prompt_learner.compound_prompts_text_parameters.1"
    assert torch.equal(avg[k].cpu(), fedavg_oracle([s[k].cpu() for s in sds])[1])
    sds[1][k][0, 0] = float("inf")
    assert not fed.check_weights_valid(sds[1])
    fed.broadcast_weights(avg)


def test_pipelined_epoch_and_eval_match_plain_loops():
    """run_epoch / test() overlap batch assembly + H2D (side stream, rotating staging) with the running step; they
    must give exactly what the plain per-batch loops give: same parameters bit for bit, same accuracy."""
    import contextlib, io
    names = synth.synthetic_classnames(10)
    items = synthetic_client_items(10, 3, seed=5, classnames=names)              # 30 images
    cfg = synth.make_cfg()
    cfg.DATALOADER = synth._NS(TRAIN_X=synth._NS(BATCH_SIZE=4), TEST=synth._NS(BATCH_SIZE=4))
    outs = []
    for pipelined in (True, False):
        dm = ClientDataManager(items[:22], [], items[20:], cfg)                  # 5 train batches, 3 test batches (4,4,2)
        t = _trainer(False)
        t.dm = dm
        with contextlib.redirect_stdout(io.StringIO()):
            if pipelined:
                t.run_epoch(0)
                acc = t.test()["accuracy"]
            else:
                t.model.train()
                for batch in dm.train_loader:
                    t.forward_backward(batch)
                t.update_lr()
                t.model.eval()
                hit = tot = 0
                for batch in dm.test_loader:
                    x, y, _ = t.parse_batch_train(batch)
                    hit += int((t.model_inference(x).argmax(1) == y).sum()); tot += int(y.numel())
                acc = 100.0 * hit / tot
        torch.cuda.synchronize()
        outs.append((t.model.engine.params.clone(), acc))
    assert torch.equal(outs[0][0], outs[1][0])
    assert outs[0][1] == outs[1][1]


def test_load_state_dict_reaches_frozen_tensors_and_failed_step_keeps_weights():
    """(medium) CustomCLIP.load_state_dict must push EVERY tensor to the kernels (frozen block weights,
    logit_scale, embeddings, prompt prefix / suffix), not only those with requires_grad; (low) a step rejected for
    a NaN image must leave the model exactly as it was (the reference raises before optim.step())."""
    img, lab = synth.make_batch(4, 10, 123)
    t = _trainer(False)
    t.model.eval()
    before = t.model(img.cuda()).clone()
    sd = {k: v.clone() for k, v in torch.nn.Module.state_dict(t.model).items()}
    sd["logit_scale"] = sd["logit_scale"] + 0.25
    k_frozen = "image_encoder.transformer.resblocks.3.mlp.c_fc.weight"
    sd[k_frozen] = (sd[k_frozen].float() * 1.05).to(sd[k_frozen].dtype)
    sd["clip_model2.visual.transformer.resblocks.3.mlp.c_fc.weight"] = sd[k_frozen]
    sd["text_encoder.positional_embedding"] = sd["text_encoder.positional_embedding"] * 0.5
    sd["clip_model2.positional_embedding"] = sd["text_encoder.positional_embedding"]
    sd["prompt_learner.token_suffix"] = sd["prompt_learner.token_suffix"].flip(0)
    t.model.load_state_dict(sd, strict=True)
    after = t.model(img.cuda()).clone()
    assert not torch.equal(before, after)
    # a fresh model built around the same tensors gives the same logits
    t2 = _trainer(False)
    t2.model.load_state_dict(sd, strict=True)
    t2.model.eval()
    assert torch.equal(t2.model(img.cuda()), after)
    fresh = t2.model.engine.logits(img.cuda(), cache_text=False)
    assert torch.equal(fresh, after)
    # ---- rejected step: parameters, momentum and the bf16 copies stay as they were
    t.model.train()
    t.forward_backward({"img": img, "label": lab})                  # one good step (momentum now non-zero)
    eng = t.model.engine
    snap = (eng.params.clone(), eng.momentum.clone(), eng.vis.w[11]["mlp.c_fc.w"].clone())
    bad = img.clone(); bad[1, 2, 3, 4] = float("nan")
    for graph in (False, True):
        t._use_graph = graph
        with pytest.raises(ValueError, match="NaN"):
            t.forward_backward({"img": bad, "label": lab})
        torch.cuda.synchronize()
        assert torch.equal(eng.params, snap[0]) and torch.equal(eng.momentum, snap[1])
        assert torch.equal(eng.vis.w[11]["mlp.c_fc.w"], snap[2])
    out = t.forward_backward({"img": img, "label": lab})            # training continues normally afterwards
    assert out["loss"] > 0 and not torch.equal(eng.params, snap[0])


def test_federated_failure_bookkeeping_and_weighted_mode():
    """Round orchestration of trainers/maple_fed.py:228-303: a client whose local training raises is dropped from the
    round (failed_clients), the average runs over the remaining clients only (divisor = their number, or their sample
    count in the weighted mode); if every client fails the round is skipped and the previous global model stays."""
    import contextlib, io
    from federated_multi_modal_b200.trainers import Datum
    C, K = 4, 3
    names = synth.synthetic_classnames(C)
    pool = synthetic_client_items(C, 3, seed=2, classnames=names)               # 12 images

    def make(bad_clients, weighted, sizes=(4, 4, 4), poison=3.0e38, drop_on_input_error=False):
        cfg = synth.make_cfg()
        cfg.FED.DROP_ON_INPUT_ERROR = drop_on_input_error
        cfg.FED.NUM_CLIENTS, cfg.FED.NUM_ROUNDS, cfg.FED.LOCAL_EPOCHS = K, 1, 1
        cfg.FED.WEIGHTED = weighted
        cfg.DATALOADER = synth._NS(TRAIN_X=synth._NS(BATCH_SIZE=2), TEST=synth._NS(BATCH_SIZE=4))
        cfg.OUTPUT_DIR = ""
        dms, off = [], 0
        for k in range(K):
            items = [Datum(impath=it.impath, label=it.label, classname=it.classname, img=it.img.clone())
                     for it in pool[off:off + sizes[k]]]
            off += sizes[k]
            if k in bad_clients:
                for it in items:
                    it.img.fill_(poison)        # finite but overflowing: "NaN/Inf in total loss" (RuntimeError) every batch
            dms.append(ClientDataManager(items, [], [], cfg))
        fed = MaPLeFederated(cfg, client_data_managers=dms, classnames=names)
        cap = {}
        orig = fed._aggregate
        def spy():
            cap["rows"] = [r.clone().cpu() for r in fed.exchange.gather()]
            cap["status"] = fed.exchange.status.cpu().clone()
            return orig()
        fed._aggregate = spy
        return fed, cap

    with contextlib.redirect_stdout(io.StringIO()):
        fed, cap = make({1}, weighted=False)
        start = fed.global_arena.clone()
        fed.train()
    assert fed.nan_stats["failed_clients"] == [1] and fed.nan_stats["total_updates"] == 1
    assert cap["status"][:, 0].tolist() == [1.0, 0.0, 1.0]
    m32, m16 = fedavg_oracle([cap["rows"][0], cap["rows"][2]])
    assert torch.equal(fed.last_mean_fp32.cpu(), m32) and torch.equal(fed.global_arena.cpu(), m16.float())
    assert not torch.equal(fed.global_arena, start)
    # the dropped client did not train (its update was rejected before the optimiser): its row is the broadcast model
    assert torch.equal(cap["rows"][1], start.cpu())
    # ---- weighted by sample count (north_star extension): 6 / 2 / 4 training images
    with contextlib.redirect_stdout(io.StringIO()):
        fed, cap = make(set(), weighted=True, sizes=(6, 2, 4))
        fed.train()
    assert cap["status"][:, 1].tolist() == [6.0, 2.0, 4.0]
    m32, _ = fedavg_oracle(cap["rows"], [6.0, 2.0, 4.0])
    assert torch.equal(fed.last_mean_fp32.cpu(), m32)
    # ---- every client fails: round skipped, global model unchanged
    with contextlib.redirect_stdout(io.StringIO()):
        fed, cap = make({0, 1, 2}, weighted=False)
        start = fed.global_arena.clone()
        fed.train()
    assert fed.nan_stats["skipped_rounds"] == 1 and fed.nan_stats["total_updates"] == 0
    assert sorted(fed.nan_stats["failed_clients"]) == [0, 1, 2]
    assert torch.equal(fed.global_arena, start)
    # ---- a NaN INPUT raises ValueError, which the reference's round loop does not catch (only RuntimeError)
    with contextlib.redirect_stdout(io.StringIO()):
        fed, cap = make({1}, weighted=False, poison=float("nan"))
        with pytest.raises(ValueError, match="NaN values in input image"):
            fed.train()
        fed, cap = make({1}, weighted=False, poison=float("nan"), drop_on_input_error=True)   # opt-in: drop instead
        fed.train()
    assert fed.nan_stats["failed_clients"] == [1] and fed.nan_stats["total_updates"] == 1
