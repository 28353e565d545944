"""Generates the committed golden fixtures by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
Outputs (small, committed): tests/golden/*.pt
  c1_fp32.pt   BASELINE config 1 (B=4, C=10, n_ctx=2, depth=9), fp32-ref oracle definition
               (SURVEY.md §8c): eval logits, train loss, features, every parameter gradient of the
               reference's own autograd (full tensors < 64k elems, strided sample + norm otherwise).
  c1_fp16_logits.pt  the reference as-is (PREC=fp16) eval logits, for information.
  c3s_fp32.pt  C=38 classes / B=2, same content (text-heavy shape of config 3).
  fedavg.pt    MaPLeFederated.safe_average_weights / check_weights_valid known answers at
               K = 2, 3, 8, 16, 17, 32, 33 incl. NaN/Inf entries.
  c2_fp32.pt   BASELINE config 2 (B=32, C=10): the shape bench.py measures. Same content as c1.
  c4s_fp32.pt  BASELINE config 4 per-client step shape (B=64, C=21). Same content.
  traj_fp32.pt 3-step training trajectories of the reference (CustomCLIP + clip_grad_norm_(1.0) +
               torch.optim.SGD(momentum 0.9, wd 5e-4), trainers/maple.py:588-598) at B=4, C=10 for the yaml's
               LR 0.0026 and for LR 0.05: per-step losses / grad norms and the final values of every
               prompt-learner tensor + strided samples of the other trainable tensors. Also the loss
               sequence of the reference AS-IS (PREC=fp16: SGD applied to fp16 parameters in place) to
               quantify the fp32-master vs fp16-in-place difference.
  edge_n4d12_fp32.pt / edge_n2d1_fp32.pt  other prompt configurations (N_CTX=4, PROMPT_DEPTH=12; PROMPT_DEPTH=1), B=2, C=3.
  ckpt_spec.pt the checkpoint dict MaPLeFederated.save_model hands to Dassl's save_checkpoint
               (trainers/maple_fed.py:367-386): target directory, top-level keys, and key/shape/dtype of
               the state_dict before and after one safe_average_weights aggregation.
Inputs are regenerated from seeds by federated_multi_modal_b200.synth on any box.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

from federated_multi_modal_b200 import synth  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402

BIG = 65536
STRIDE = 97


def pack_grad(g: torch.Tensor):
    g = g.detach().float()
    if g.numel() <= BIG:
        return dict(full=g.clone())
    flat = g.reshape(-1)
    return dict(sample=flat[::STRIDE].clone(), stride=STRIDE, norm=flat.double().norm().item(),
                sum=flat.double().sum().item(), shape=tuple(g.shape))


def run_case(B, C, seed_clip=0, seed_pl=1, seed_batch=123, fp32=True, n_ctx=2, depth=9):
    torch.manual_seed(0)
    cfg = synth.make_cfg(n_ctx=n_ctx, depth=depth)
    names = synth.synthetic_classnames(C)
    sd = synth.random_clip_state_dict(seed_clip)
    pl = synth.random_prompt_learner_state(seed_pl, n_ctx=n_ctx, depth=depth)
    model = rh.build_reference_customclip(sd, names, cfg, pl, fp32=fp32)
    img, lab = synth.make_batch(B, C, seed_batch)
    out = dict(meta=dict(B=B, C=C, seed_clip=seed_clip, seed_pl=seed_pl, seed_batch=seed_batch,
                         n_ctx=n_ctx, depth=depth, fp32=fp32, torch=torch.__version__))
    model.eval()
    feats = {}
    with torch.no_grad():
        out["logits_eval"] = model(img if fp32 else img.half()).float().clone()
    if not fp32:
        return out, model
    # features through the reference's own sub-modules
    with torch.no_grad():
        prompts, shared, dtx, dvs = model.prompt_learner()
        out["text_features"] = model.text_encoder(prompts, model.tokenized_prompts, dtx).clone()
        out["image_features"] = model.image_encoder(img, shared, dvs, None).clone()
        out["shared_ctx"] = shared.clone()
    # per-block activations via forward hooks on the reference modules (batch-major)
    acts = {}
    hooks = []
    for tower, blocks in (("vis", model.image_encoder.transformer.resblocks),
                          ("txt", model.text_encoder.transformer.resblocks)):
        for li in (0, 1, 8, 11):
            def mk(name):  # noqa: E306
                def hook(mod, inp, outp):
                    x = outp[0].detach().permute(1, 0, 2)  # LND -> NLD
                    acts[name] = x[:2, ::16, ::8].clone()  # small strided sample
                return hook
            hooks.append(blocks[li].register_forward_hook(mk(f"{tower}{li}")))
    model.train()
    loss = model(img, lab)
    for h in hooks:
        h.remove()
    model.zero_grad()
    loss.backward()
    out["loss"] = loss.detach().clone()
    out["acts"] = acts
    grads = {}
    for n, p in model.named_parameters():
        if p.requires_grad and p.grad is not None:
            grads[n] = pack_grad(p.grad)
    out["grads"] = grads
    out["n_trainable"] = sum(1 for p in model.parameters() if p.requires_grad)
    out["n_with_grad"] = len(grads)
    out["n_state_dict_keys"] = len(model.state_dict())
    return out, model


def fedavg_cases():
    _, _, fed = rh.load_reference()
    g = torch.Generator().manual_seed(7)
    cases = {}
    for K in (2, 3, 8, 16, 17, 32, 33):
        dicts = []
        for k in range(K):
            a = torch.randn(1000, generator=g)
            b = (torch.randn(33, 7, generator=g) * 3).half()
            if k == 1:
                a[3] = float("nan"); a[4] = float("inf"); a[5] = float("-inf")
            dicts.append({"a": a, "b": b, "s": torch.tensor(2.6592600 + 0.001 * k)})
        avg = fed.MaPLeFederated.safe_average_weights(None, dicts, K)
        cases[K] = dict(inputs=dicts, out={k: v.clone() for k, v in avg.items()},
                        fp32_mean={k: torch.nan_to_num(torch.stack([d[k].float() for d in dicts]), nan=0.0,
                                                       posinf=1e4, neginf=-1e4).mean(0) for k in dicts[0]},
                        valid=[bool(fed.MaPLeFederated.check_weights_valid(None, d)) for d in dicts])
    return cases


def trajectory(lr, steps=3, B=4, C=10, fp32=True, seed_batch=500):
    """The reference's own step (trainers/maple.py:588-598) repeated: model(image, label) -> zero_grad ->
    backward -> clip_grad_norm_(model.parameters(), 1.0) -> optim.step(), Dassl's SGD defaults."""
    torch.manual_seed(0)
    cfg = synth.make_cfg()
    model = rh.build_reference_customclip(synth.random_clip_state_dict(0), synth.synthetic_classnames(C), cfg,
                                          synth.random_prompt_learner_state(1), fp32=fp32)
    model.train()
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.SGD(params, lr=lr, momentum=0.9, weight_decay=5e-4, dampening=0, nesterov=False)
    init = {n: p.detach().float().clone() for n, p in model.named_parameters() if p.requires_grad}
    losses, norms = [], []
    for s in range(steps):
        img, lab = synth.make_batch(B, C, seed_batch + s)
        loss = model(img if fp32 else img.half(), lab)
        opt.zero_grad()
        loss.backward()
        norms.append(float(torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0, error_if_nonfinite=False)))
        opt.step()
        losses.append(float(loss))
    out = dict(meta=dict(B=B, C=C, lr=lr, steps=steps, seed_batch=seed_batch, fp32=fp32, momentum=0.9,
                         weight_decay=5e-4), losses=losses, grad_norms=norms)
    if fp32:
        model.eval()
        with torch.no_grad():
            img, _ = synth.make_batch(B, C, seed_batch + steps)
            out["logits_after"] = model(img).float().clone()
        out["final"], out["delta"] = {}, {}
        for n, p in model.named_parameters():
            if p.requires_grad and n in init:
                out["final"][n] = pack_grad(p)          # full below 64k elements, strided sample + norm above
                out["delta"][n] = pack_grad(p.detach().float() - init[n])
    return out


def checkpoint_spec():
    """What MaPLeFederated.save_model builds (trainers/maple_fed.py:367-386), captured by swapping the (stubbed)
    Dassl save_checkpoint of the imported reference module for a recorder."""
    import types
    _, _, fed = rh.load_reference()
    cfg = synth.make_cfg()
    m = rh.build_reference_customclip(synth.random_clip_state_dict(0), synth.synthetic_classnames(10), cfg,
                                      synth.random_prompt_learner_state(1), fp32=False)
    sd0 = m.state_dict()
    averaged = fed.MaPLeFederated.safe_average_weights(None, [sd0, sd0], [0, 1])
    got = {}
    made = []
    saved = (fed.save_checkpoint, fed.mkdir_if_missing)
    fed.save_checkpoint = lambda state, d, is_best=False, **kw: got.update(state=state, dir=d, is_best=is_best)
    fed.mkdir_if_missing = lambda d: made.append(d)
    try:
        fake = types.SimpleNamespace(
            cfg=types.SimpleNamespace(OUTPUT_DIR="OUT", VERBOSE=False, OPTIM=types.SimpleNamespace(MAX_EPOCH=2),
                                      dump=lambda: "cfg-dump"),
            global_weights=averaged)
        fed.MaPLeFederated.save_model(fake, directory="OUT")
    finally:
        fed.save_checkpoint, fed.mkdir_if_missing = saved
    st = got["state"]
    spec = lambda d: [(k, tuple(v.shape), str(v.dtype)) for k, v in d.items()]
    return dict(target_dir=got["dir"], made_dirs=made, top_keys=list(st.keys()), epoch=st["epoch"],
                optimizer=st["optimizer"], scheduler=st["scheduler"], cfg=st["cfg"],
                spec_before_aggregation=spec(sd0), spec_after_aggregation=spec(st["state_dict"]),
                logit_scale_after=st["state_dict"]["logit_scale"].clone(),
                ctx_after=st["state_dict"]["prompt_learner.ctx"].clone())


def edge_cases():
    """Other prompt configurations of the reference (cfg.TRAINER.MAPLE.N_CTX / PROMPT_DEPTH): the deepest and widest
    the CTX_INIT path allows (n_ctx=4, depth=12: prompts spliced into layers 1..11, T_v = 201) and the shallowest
    (depth=1: no deep prompts, no compound projections). B=2, C=3."""
    torch.set_num_threads(8)
    for tag, n_ctx, depth in (("n4d12", 4, 12), ("n2d1", 2, 1)):
        o, _ = run_case(2, 3, seed_batch=77, n_ctx=n_ctx, depth=depth)
        torch.save(o, os.path.join(HERE, f"edge_{tag}_fp32.pt"))
        print(tag, "loss", o["loss"].item(), "grads", o["n_with_grad"], "/", o["n_trainable"])


if __name__ == "__main__" and "--edge-only" in sys.argv:
    import contextlib, io
    with contextlib.redirect_stderr(io.StringIO()):
        edge_cases()
    sys.exit(0)


def extra_cases():
    torch.set_num_threads(8)
    o2, _ = run_case(32, 10, seed_batch=2032)
    torch.save(o2, os.path.join(HERE, "c2_fp32.pt"))
    print("c2 loss", o2["loss"].item())
    o4, _ = run_case(64, 21, seed_batch=4064)
    torch.save(o4, os.path.join(HERE, "c4s_fp32.pt"))
    print("c4s loss", o4["loss"].item())
    tr = {"lr0.0026": trajectory(0.0026), "lr0.05": trajectory(0.05),
          "lr0.0026_fp16_as_is": trajectory(0.0026, fp32=False), "lr0.05_fp16_as_is": trajectory(0.05, fp32=False)}
    torch.save(tr, os.path.join(HERE, "traj_fp32.pt"))
    for k, v in tr.items():
        print(k, "losses", v["losses"], "norms", v["grad_norms"])
    torch.save(checkpoint_spec(), os.path.join(HERE, "ckpt_spec.pt"))


if __name__ == "__main__" and "--extra-only" in sys.argv:
    import contextlib, io
    with contextlib.redirect_stderr(io.StringIO()):
        extra_cases()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".pt"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
    sys.exit(0)


if __name__ == "__main__":
    import contextlib, io
    torch.set_num_threads(8)
    extra_cases()
    edge_cases()
    o, _ = run_case(4, 10)
    torch.save(o, os.path.join(HERE, "c1_fp32.pt"))
    print("c1 loss", o["loss"].item(), "grads", o["n_with_grad"], "/", o["n_trainable"])
    o16, _ = run_case(4, 10, fp32=False)
    torch.save(o16, os.path.join(HERE, "c1_fp16_logits.pt"))
    rel = ((o16["logits_eval"] - o["logits_eval"]).abs().max() / o["logits_eval"].abs().max()).item()
    print("fp16-ref vs fp32-ref max rel logit err", rel)
    o3, _ = run_case(2, 38, seed_batch=321)
    torch.save(o3, os.path.join(HERE, "c3s_fp32.pt"))
    print("c3s loss", o3["loss"].item())
    with contextlib.redirect_stdout(io.StringIO()):
        fc = fedavg_cases()
    torch.save(fc, os.path.join(HERE, "fedavg.pt"))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".pt"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


def state_dict_spec():
    """Key / shape / dtype list of the reference CustomCLIP.state_dict() in its default fp16 mode, and the
    names its freeze policy leaves trainable -> tests/golden/state_dict_spec.pt (drop-in compatibility pin)."""
    cfg = synth.make_cfg()
    m = rh.build_reference_customclip(synth.random_clip_state_dict(0), synth.synthetic_classnames(10), cfg,
                                      synth.random_prompt_learner_state(1), fp32=False)
    spec = [(k, tuple(v.shape), str(v.dtype)) for k, v in m.state_dict().items()]
    train = sorted(n for n, p in m.named_parameters() if p.requires_grad)
    torch.save(dict(spec=spec, trainable=train), os.path.join(HERE, "state_dict_spec.pt"))


if __name__ == "__main__":
    state_dict_spec()
