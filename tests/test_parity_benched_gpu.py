"""GPU parity at the configurations that are BENCHED, through the path that is benched.

* config 2 (B=32, C=10) and the config-4 client step (B=64, C=21): the CUDA-graph-replayed
  ``MaPLe.forward_backward`` (split-K GEMM tails, multi-unit attention CTAs, fused clip+SGD) against the
  reference's own outputs (``tests/golden/c2_fp32.pt`` / ``c4s_fp32.pt``: loss, logits, all 145 gradients of its
  autograd) — logits within 2e-2 of max |logit| (north_star), every gradient tensor reported and bounded.
* 3-step training trajectories of the reference (CustomCLIP + ``clip_grad_norm_(1.0)`` + ``torch.optim.SGD``,
  trainers/maple.py:588-598; ``tests/golden/traj_fp32.pt``) against ``MaPLe.forward_backward`` x 3: per-step
  loss and pre-clip gradient norm, final prompt-learner tensors, logits after the updates.
  The reference (fp16 mode) applies SGD to its fp16 parameters in place; this implementation updates fp32
  master copies. The fixture also holds the loss sequence of the reference as-is (fp16), and the test prints
  both distances: the fp32-master trajectory sits closer to the fp32 reference than the reference's own fp16
  mode does.
* evaluation after graph-replayed training steps must not reuse stale text features.

Each test appends its measured numbers to ``gpurun_out/parity_report.jsonl`` (copied to ``profiles/`` by hand).
"""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from helpers import REPO, customclip_state_dict, load_golden
from federated_multi_modal_b200 import synth

if torch.cuda.is_available():
    from federated_multi_modal_b200.trainers import MaPLe

F32 = torch.float32


def _report(rec):
    d = os.environ.get("MFK_REPORT_DIR", os.path.join(REPO, "gpurun_out"))
    try:
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "parity_report.jsonl"), "a") as f:
            f.write(json.dumps(rec) + "\n")
    except OSError:
        pass


def _rel(a, b):
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


def _trainer(C, graph=True, lr=None):
    """MaPLe trainer whose engine holds exactly the parameter values the fixtures were generated with (the
    reference's fp32-ref definition keeps ``ctx`` in fp32; the module's fp16 ``ctx`` would add a rounding the
    fixture does not have)."""
    cfg = synth.make_cfg()
    cfg.USE_CUDA_GRAPH = graph
    t = MaPLe(cfg, client_id=0, classnames=synth.synthetic_classnames(C))
    t.model.prompt_learner.load_state_dict(synth.random_prompt_learner_state(1), strict=False)
    t.model.load_state_dict(torch.nn.Module.state_dict(t.model))
    sd, _ = customclip_state_dict(C)
    eng = t.model.engine
    eng.p["prompt_learner.ctx"].copy_(sd["prompt_learner.ctx"])
    eng.repack_trainable()
    t.model._arena_newer = True
    if lr is not None:
        t.optim.lr = lr
    t.model.train()
    return t


def _grad_table(eng, golden_grads, coef):
    """Per-tensor comparison of the engine's gradient arena (clipped in place by ``coef``) with the reference's
    autograd gradients: cosine, max error relative to the tensor's max, norm ratio."""
    rows = {}
    for name, packed in golden_grads.items():
        g = eng.g[name].detach().float().cpu() / coef
        if "full" in packed:
            ref, got = packed["full"], g
            nr = g.double().norm().item() / max(ref.double().norm().item(), 1e-30)
        else:
            ref, got = packed["sample"], g.reshape(-1)[::packed["stride"]]
            nr = g.double().norm().item() / max(packed["norm"], 1e-30)
        cos = torch.nn.functional.cosine_similarity(got.reshape(-1).double(), ref.reshape(-1).double(), dim=0).item()
        rows[name] = (cos, _rel(got, ref), nr)
    return rows


@pytest.mark.parametrize("fixture", ["c2_fp32.pt", "c4s_fp32.pt"])
def test_graph_replayed_step_vs_reference_at_benched_shapes(fixture):
    G = load_golden(fixture)
    m = G["meta"]
    B, C = m["B"], m["C"]
    t = _trainer(C, graph=True)
    eng = t.model.engine
    img, lab = synth.make_batch(B, C, m["seed_batch"])
    out = t.forward_backward({"img": img.pin_memory(), "label": lab.pin_memory()})
    assert t._graph is not None and t._graph_B == B                       # the replayed graph produced these numbers
    assert eng.vis.gemm_ws is not None and eng.vis.M == B * 199           # split-K workspace of the vision GEMM tails
    loss_ref = G["loss"].item()
    assert abs(out["loss"] - loss_ref) < 2e-2 * abs(loss_ref), (out["loss"], loss_ref)
    logits = eng._buf("head.logits", (B, C), F32).cpu()
    e_logit = _rel(logits, G["logits_eval"])
    assert e_logit < 2e-2, e_logit
    top1 = (logits.argmax(1) == G["logits_eval"].argmax(1)).float().mean().item()
    _, norm, _ = t.read_step_result()
    ref_norm = sum(((p["full"].double().norm().item() if "full" in p else p["norm"]) ** 2) for p in G["grads"].values()) ** 0.5
    assert abs(norm - ref_norm) < 3e-2 * ref_norm, (norm, ref_norm)       # pre-clip total gradient norm
    coef = min(1.0, 1.0 / (norm + 1e-6))
    assert set(G["grads"]) <= set(eng.g)
    tab = _grad_table(eng, G["grads"], coef)
    worst_cos = sorted(tab.items(), key=lambda kv: kv[1][0])[:5]
    worst_rel = sorted(tab.items(), key=lambda kv: -kv[1][1])[:5]
    worst_nr = sorted(tab.items(), key=lambda kv: -abs(kv[1][2] - 1))[:5]
    print(f"{fixture}: loss {out['loss']:.5f} (ref {loss_ref:.5f}), logits rel err {e_logit:.2e}, raw top-1 agreement "
          f"{top1:.3f}, grad norm {norm:.4f} (ref {ref_norm:.4f})")
    print("  lowest cosine:", [(k, round(v[0], 5)) for k, v in worst_cos])
    print("  largest max-rel:", [(k, round(v[1], 4)) for k, v in worst_rel])
    print("  norm ratio furthest from 1:", [(k, round(v[2], 4)) for k, v in worst_nr])
    _report({"test": "graph_step_vs_reference", "fixture": fixture, "B": B, "C": C, "loss": out["loss"],
             "loss_ref": loss_ref, "logits_rel_err": e_logit, "raw_top1_agreement": top1, "grad_norm": norm,
             "grad_norm_ref": ref_norm, "min_cos": worst_cos[0][1][0], "max_rel": worst_rel[0][1][1],
             "worst_norm_ratio": worst_nr[0][1][2],
             "per_tensor": {k: [round(x, 6) for x in v] for k, v in tab.items()}})
    # bf16 tensor-core backward against fp32 autograd: direction, magnitude and elementwise error of EVERY tensor
    # (measured: cosine >= 0.9996, max error <= 4.8 % of the tensor's max, norms within 1.7 %)
    assert worst_cos[0][1][0] > 0.999, worst_cos
    assert worst_rel[0][1][1] < 0.07, worst_rel
    assert abs(worst_nr[0][1][2] - 1) < 0.03, worst_nr


@pytest.mark.parametrize("key", ["lr0.0026", "lr0.05"])
def test_three_step_trajectory_vs_reference(key):
    TR = load_golden("traj_fp32.pt")
    G, G16 = TR[key], TR[key + "_fp16_as_is"]
    m = G["meta"]
    t = _trainer(m["C"], graph=True, lr=m["lr"])
    eng = t.model.engine
    assert (t.optim.momentum, t.optim.weight_decay) == (m["momentum"], m["weight_decay"])
    init = {k: v.clone() for k, v in eng.p.items()}
    losses, norms = [], []
    for s in range(m["steps"]):
        img, lab = synth.make_batch(m["B"], m["C"], m["seed_batch"] + s)
        t.optim.lr = m["lr"]
        losses.append(t.forward_backward({"img": img.pin_memory(), "label": lab.pin_memory()})["loss"])
        norms.append(t.read_step_result()[1])
    d_loss = max(abs(a - b) / abs(b) for a, b in zip(losses, G["losses"]))
    d_norm = max(abs(a - b) / abs(b) for a, b in zip(norms, G["grad_norms"]))
    d_loss16 = max(abs(a - b) / abs(b) for a, b in zip(G16["losses"], G["losses"]))
    d_norm16 = max(abs(a - b) / abs(b) for a, b in zip(G16["grad_norms"], G["grad_norms"]))
    print(f"{key}: losses {losses} vs ref {G['losses']} (max rel {d_loss:.2e}; reference fp16-as-is vs fp32: {d_loss16:.2e})")
    print(f"{key}: grad norms {norms} vs ref {G['grad_norms']} (max rel {d_norm:.2e}; fp16-as-is: {d_norm16:.2e})")
    assert d_loss < 1e-2 and d_norm < 3e-2
    # parameter displacement after 3 clipped SGD+momentum steps: compare the UPDATE (final - initial), which is what
    # the optimiser produced, tensor by tensor
    tab = {}
    for name, packed in G["delta"].items():
        if name not in eng.p or "proj_vis_to_lang" in name:
            continue
        d = (eng.p[name] - init[name]).float().cpu()
        ref = packed["full"] if "full" in packed else packed["sample"]
        got = d if "full" in packed else d.reshape(-1)[::packed["stride"]]
        if ref.abs().max().item() == 0.0:
            assert got.abs().max().item() == 0.0, name
            continue
        cos = torch.nn.functional.cosine_similarity(got.reshape(-1).double(), ref.reshape(-1).double(), dim=0).item()
        tab[name] = (cos, _rel(got, ref))
    pl = {k: v for k, v in tab.items() if k.startswith("prompt_learner.")}
    assert len(pl) >= 20
    worst_cos = sorted(tab.items(), key=lambda kv: kv[1][0])[:5]
    worst_rel = sorted(tab.items(), key=lambda kv: -kv[1][1])[:5]
    print("  update: lowest cosine", [(k, round(v[0], 5)) for k, v in worst_cos])
    print("  update: largest max-rel", [(k, round(v[1], 4)) for k, v in worst_rel])
    # final values of the prompt learner (what FedAvg ships) relative to their own scale
    e_final = {}
    for name, packed in G["final"].items():
        if name in eng.p and name.startswith("prompt_learner.") and "proj_vis_to_lang" not in name:
            got = eng.p[name].float().cpu()
            e_final[name] = _rel(got, packed["full"]) if "full" in packed else \
                _rel(got.reshape(-1)[::packed["stride"]], packed["sample"])
    t.model.eval()
    img, _ = synth.make_batch(m["B"], m["C"], m["seed_batch"] + m["steps"])
    lg = t.model(img.cuda()).cpu()
    e_logit = _rel(lg, G["logits_after"])
    # forward precision at the TRAINED parameters, isolated from trajectory divergence: the engine's fp32 mode (pinned
    # to the reference at 2e-5 by test_fp32_mode_logits_vs_reference_golden) evaluated on the same arena
    e_same = _rel(lg, eng.logits(img.cuda(), precision="fp32").cpu())
    print(f"  logits after {m['steps']} steps: rel err vs reference {e_logit:.2e}, vs fp32 mode at the same parameters "
          f"{e_same:.2e}; worst final prompt tensor rel err {max(e_final.values()):.2e}")
    _report({"test": "trajectory", "key": key, "losses": losses, "losses_ref": G["losses"], "grad_norms": norms,
             "grad_norms_ref": G["grad_norms"], "loss_rel": d_loss, "norm_rel": d_norm,
             "reference_fp16_as_is_loss_rel": d_loss16, "reference_fp16_as_is_norm_rel": d_norm16,
             "update_min_cos": worst_cos[0][1][0], "update_max_rel": worst_rel[0][1][1],
             "final_prompt_max_rel": max(e_final.values()), "logits_after_rel": e_logit,
             "logits_after_rel_same_params_fp32_mode": e_same})
    assert worst_cos[0][1][0] > 0.99, worst_cos
    assert worst_rel[0][1][1] < 0.1, worst_rel
    # final tensors: error of the update (<= 5 % of it) relative to the tensor (measured 3e-4 at lr 0.0026, 4e-3 at 0.05)
    assert max(e_final.values()) < 1e-2, sorted(e_final.items(), key=lambda kv: -kv[1])[:3]
    # north_star tolerance 2e-2, at the trained parameters: against fp32 arithmetic on the same arena, and against the
    # reference's own trained model (measured 3e-3 / 4e-3). History: with round-to-nearest bf16 copies of the trainable
    # block weights this was 2.1e-2 / 4.5e-2 at lr 0.0026 — sub-ulp updates of fp16-grid weights over-shoot coherently
    # (see repack_grouped_kernel); the CPU oracle with RNE bf16 operands shows the same 5e-2, its fp32 form matches the
    # reference's trajectory to 1.4e-5.
    assert e_same < 1e-2 and e_logit < 1e-2


def test_eval_after_graph_replayed_steps_uses_fresh_text_features():
    """Graph replay runs no Python inside forward_backward, so the text-feature cache must be invalidated by the
    replay itself: evaluate, train two more replayed steps, evaluate again — the second evaluation must equal an
    uncached one (and differ from the first)."""
    C, B = 10, 4
    t = _trainer(C, graph=True, lr=0.05)
    eng = t.model.engine
    batches = [synth.make_batch(B, C, 900 + i) for i in range(4)]
    probe = synth.make_batch(B, C, 77)[0].cuda()
    for i in range(2):
        t.optim.lr = 0.05
        t.forward_backward({"img": batches[i][0], "label": batches[i][1]})
    assert t._graph is not None
    t.model.eval()
    a = t.model(probe).clone()
    t.model.train()
    for i in range(2, 4):
        t.optim.lr = 0.05
        t.forward_backward({"img": batches[i][0], "label": batches[i][1]})
    t.model.eval()
    b = t.model(probe).clone()
    fresh = eng.logits(probe, cache_text=False).clone()
    assert torch.equal(b, fresh)
    assert not torch.equal(a, b)
