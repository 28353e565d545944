"""GPU: numerics of the standalone hook surface and the checkpoint round trip.

north_star keeps the reference's hooks callable on their own: ``MultiModalPromptLearner.forward``
(trainers/maple.py:177-218), ``TextEncoder.forward`` (52-79), ``VisionTransformer_MaPLe.forward``
(clip/model.py:509-572) and ``ResidualAttentionBlock_MaPLe.forward([x, deep, counter])`` (307-352). They are
executed here and compared with what the unmodified reference's own sub-modules produced on the same seeds
(``tests/golden/c1_fp32.pt``: text_features, image_features, shared_ctx, per-block activation samples).
"""
import contextlib
import io
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from helpers import load_golden
from federated_multi_modal_b200 import synth

if torch.cuda.is_available():
    from federated_multi_modal_b200.clip import build_model
    from federated_multi_modal_b200.trainers import ClientDataManager, CustomCLIP, MaPLeFederated
    from federated_multi_modal_b200.trainers.client_datamanager import synthetic_client_items

DD = {"trainer": "MaPLe", "vision_depth": 0, "language_depth": 0, "vision_ctx": 0, "language_ctx": 0,
      "maple_length": 2}


def _rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


def _custom_clip(C=10):
    clip = build_model(synth.random_clip_state_dict(0), DD)
    m = CustomCLIP(synth.make_cfg(), synth.synthetic_classnames(C), clip)
    m.prompt_learner.load_state_dict(synth.random_prompt_learner_state(1), strict=False)
    return m.cuda().eval()


def test_hook_forwards_vs_reference_submodules():
    G = load_golden("c1_fp32.pt")
    meta = G["meta"]
    img, _ = synth.make_batch(meta["B"], meta["C"], meta["seed_batch"])
    m = _custom_clip(meta["C"])
    # ---- MultiModalPromptLearner.forward
    prompts, shared, deep_text, deep_vis = m.prompt_learner()
    assert prompts.shape == (meta["C"], 77, 512) and shared.shape == (2, 768)
    assert len(deep_text) == 8 and len(deep_vis) == 8
    assert all(t.shape == (2, 512) for t in deep_text) and all(v.shape == (2, 768) for v in deep_vis)
    e_shared = _rel(shared, G["shared_ctx"])
    assert e_shared < 2e-3, e_shared                      # fp16 ctx / fp16 output against the fp32 reference
    # even projections map text prompts to vision prompts with the fp32 Linear of the module
    pl = m.prompt_learner
    for i in (0, 2, 4, 6):
        want = torch.nn.functional.linear(pl.compound_prompts_text_parameters[i // 2].float(),
                                          pl.compound_prompt_projections[i].weight.float(),
                                          pl.compound_prompt_projections[i].bias.float())
        assert _rel(deep_vis[i], want) < 1e-5
        assert deep_text[i] is pl.compound_prompts_text_parameters[i // 2]
    # ---- per-block activations via forward hooks on OUR modules: every ResidualAttentionBlock_MaPLe.forward runs
    acts, hooks = {}, []
    for tower, blocks in (("vis", m.image_encoder.transformer.resblocks), ("txt", m.text_encoder.transformer.resblocks)):
        for li in (0, 1, 8, 11):
            def mk(name):
                def hook(mod, inp, outp):
                    assert isinstance(outp, list) and len(outp) == 3
                    acts[name] = outp[0].detach().permute(1, 0, 2)[:2, ::16, ::8].float().cpu().clone()
                    acts[name + ".counter"] = outp[2]
                return hook
            hooks.append(blocks[li].register_forward_hook(mk(f"{tower}{li}")))
    # ---- TextEncoder.forward / VisionTransformer_MaPLe.forward
    tf = m.text_encoder(prompts, m.tokenized_prompts, deep_text)
    imf = m.image_encoder(img.cuda(), shared, deep_vis)
    for h in hooks:
        h.remove()
    assert tf.shape == (meta["C"], 512) and imf.shape == (meta["B"], 512)
    e_tf, e_if = _rel(tf, G["text_features"]), _rel(imf, G["image_features"])
    print(f"hook forwards vs reference: shared_ctx {e_shared:.2e}, text_features {e_tf:.2e}, image_features {e_if:.2e}")
    assert e_tf < 2e-2 and e_if < 2e-2
    errs = {k: _rel(v, G["acts"][k]) for k, v in acts.items() if not k.endswith(".counter")}
    print("  per-block activation samples:", {k: f"{v:.1e}" for k, v in errs.items()})
    assert set(errs) == set(G["acts"]) and max(errs.values()) < 2e-2, errs
    # the splice counter advances once per spliced layer (layers 1..8), as in clip/model.py:320-349
    assert [acts[f"vis{l}.counter"] for l in (0, 1, 8, 11)] == [0, 1, 8, 8]
    assert [acts[f"txt{l}.counter"] for l in (0, 1, 8, 11)] == [0, 1, 8, 8]
    # logits from the hook outputs == CustomCLIP.forward's eval logits at bf16 tolerance
    tn = torch.nn.functional.normalize(tf.float(), dim=-1)
    im = torch.nn.functional.normalize(imf.float(), dim=-1)
    lg = m.logit_scale.exp().clamp(max=100).float() * im @ tn.t()
    assert _rel(lg, G["logits_eval"]) < 2e-2


def test_single_block_forward_list_protocol():
    """ResidualAttentionBlock_MaPLe.forward([x(L,N,D), deep, counter]) on its own: layer 0 never splices, layer 3
    replaces the LAST n rows (vision) / rows 1..n (text) with deep[counter] and returns counter + 1."""
    m = _custom_clip(4)
    g = torch.Generator().manual_seed(3)
    n = 2
    for tower, blocks, D, T in (("vis", m.image_encoder.transformer.resblocks, 768, 199),
                                ("txt", m.text_encoder.transformer.resblocks, 512, 77)):
        x = torch.randn(T, 3, D, generator=g).cuda()
        deep = [torch.randn(n, D, generator=g).cuda() * 5 for _ in range(8)]
        o0 = blocks[0]([x, deep, 0])
        assert o0[2] == 0 and o0[0].shape == x.shape and o0[1] is deep
        o3 = blocks[3]([x, deep, 2])
        assert o3[2] == 3
        xs = x.clone()
        spliced = deep[2].half().float()[:, None, :]     # the reference splices `.half()` prompts (clip/model.py:327,344)
        if tower == "vis":
            xs[T - n:] = spliced
        else:
            xs[1:1 + n] = spliced
        o3b = blocks[3]([xs, [], 0])           # pre-spliced input, no prompts left: same result, counter untouched
        assert o3b[2] == 0 and torch.equal(o3[0], o3b[0])
        assert not torch.equal(o3[0], blocks[3]([x, [], 0])[0])
        # numerics of one block against plain torch fp32 on the same module parameters
        blk = blocks[3]
        xf = xs.float()
        h = torch.nn.functional.layer_norm(xf, (D,), blk.ln_1.weight.float(), blk.ln_1.bias.float(), 1e-5)
        mask = blk.attn_mask.to(xf.device).float() if blk.attn_mask is not None else None
        a = torch.nn.functional.multi_head_attention_forward(
            h, h, h, D, blk.n_head, blk.attn.in_proj_weight.float(), blk.attn.in_proj_bias.float(), None, None, False,
            0.0, blk.attn.out_proj.weight.float(), blk.attn.out_proj.bias.float(), training=False, need_weights=False,
            attn_mask=mask)[0]
        x2 = xf + a
        h2 = torch.nn.functional.layer_norm(x2, (D,), blk.ln_2.weight.float(), blk.ln_2.bias.float(), 1e-5)
        u = torch.nn.functional.linear(h2, blk.mlp.c_fc.weight.float(), blk.mlp.c_fc.bias.float())
        want = x2 + torch.nn.functional.linear(u * torch.sigmoid(1.702 * u), blk.mlp.c_proj.weight.float(),
                                               blk.mlp.c_proj.bias.float())
        assert _rel(o3[0], want) < 2e-2, (tower, _rel(o3[0], want))


def test_federated_checkpoint_round_trip(tmp_path):
    """train one round -> finalize_training saves ``MultiModalPromptLearner_Aggregator/model.pth.tar-<MAX_EPOCH>`` in
    the reference's layout (every tensor fp16 after an aggregation) -> a second aggregator loads it
    (load_model -> check_weights_valid -> broadcast_weights, trainers/maple_fed.py:388-411): bit-equal trainable
    arena on every client, momentum dropped, identical logits."""
    C, K = 4, 2
    spec = load_golden("ckpt_spec.pt")

    def make(out_dir):
        cfg = synth.make_cfg()
        cfg.FED.NUM_CLIENTS, cfg.FED.NUM_ROUNDS, cfg.FED.LOCAL_EPOCHS = K, 1, 1
        cfg.DATALOADER = synth._NS(TRAIN_X=synth._NS(BATCH_SIZE=2), TEST=synth._NS(BATCH_SIZE=4))
        cfg.OUTPUT_DIR = out_dir
        cfg.dump = lambda: "cfg-dump"
        names = synth.synthetic_classnames(C)
        pool = synthetic_client_items(C, 2, seed=0, classnames=names)
        dms = [ClientDataManager(pool[k::K], [], pool[:4], cfg) for k in range(K)]
        return MaPLeFederated(cfg, client_data_managers=dms, classnames=names)

    with contextlib.redirect_stdout(io.StringIO()):
        fed = make(str(tmp_path))
        fed.train()
    path = os.path.join(str(tmp_path), "MultiModalPromptLearner_Aggregator", "model.pth.tar-2")
    assert os.path.exists(path)
    ck = torch.load(path, map_location="cpu", weights_only=False)
    assert list(ck.keys()) == spec["top_keys"]
    ref_layout = {k: (s, d) for k, s, d in spec["spec_after_aggregation"]}
    for k, v in ck["state_dict"].items():
        s, d = ref_layout[k]
        assert str(v.dtype) == d == "torch.float16", k
        if "token_" not in k:                 # prefix / suffix depend on the class count (4 here, 10 in the fixture)
            assert tuple(v.shape) == s, k
    assert set(ck["state_dict"]) == set(ref_layout)
    e0 = fed.clients[0].model.engine
    n = e0.n_update
    # the saved tensors are the fp16-cast global model
    assert torch.equal(ck["state_dict"]["prompt_learner.compound_prompts_text_parameters.0"].float(),
                       e0.p["prompt_learner.compound_prompts_text_parameters.0"].cpu())
    img = synth.make_batch(4, C, 5)[0].cuda()
    want_logits = e0.logits(img, cache_text=False).clone()
    with contextlib.redirect_stdout(io.StringIO()):
        fed2 = make("")
        before = fed2.clients[0].model.engine.params[:n].clone()
        # give the second aggregator's clients some optimiser state to lose
        for t in fed2.clients:
            t.model.engine.momentum.fill_(1.0); t.model.engine.mom_initialized = True
        fed2.load_model(str(tmp_path), epoch=2)
    assert not torch.equal(before, e0.params[:n])
    for t in fed2.clients:
        eng = t.model.engine
        assert torch.equal(eng.params[:n], e0.params[:n])
        assert not eng.mom_initialized and float(eng.momentum.abs().max()) == 0.0
        assert torch.equal(eng.logit_scale, e0.logit_scale)
    assert torch.equal(fed2.global_arena, e0.params[:n])
    got_logits = fed2.clients[0].model.engine.logits(img, cache_text=False)
    assert torch.equal(got_logits, want_logits)
    with pytest.raises(FileNotFoundError):
        fed2.load_model(str(tmp_path), epoch=3)
