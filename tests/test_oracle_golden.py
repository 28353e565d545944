"""CPU: pins oracle/maple_cpu.py to the golden vectors generated from the unmodified reference."""
import pytest
import torch

from oracle.maple_cpu import MapleOracle, fedavg_oracle
from federated_multi_modal_b200 import synth
from helpers import load_golden, customclip_state_dict, check_grad_against_golden


@pytest.mark.parametrize("fixture", ["c1_fp32.pt", "c3s_fp32.pt", "c2_fp32.pt", "edge_n4d12_fp32.pt", "edge_n2d1_fp32.pt"])
def test_oracle_matches_reference_autograd(fixture):
    torch.set_num_threads(8)
    G = load_golden(fixture)
    m = G["meta"]
    sd, tok = customclip_state_dict(m["C"], m["seed_clip"], m["seed_pl"], n_ctx=m["n_ctx"], depth=m["depth"])
    orc = MapleOracle(sd, tok, n_ctx=m["n_ctx"], depth=m["depth"])
    img, lab = synth.make_batch(m["B"], m["C"], m["seed_batch"])
    out = orc.forward_backward(img, lab)
    # fp32 re-association only: tight tolerances
    assert torch.allclose(out["logits"], G["logits_eval"], rtol=0, atol=2e-5)
    assert torch.allclose(out["image_features"], G["image_features"], rtol=0, atol=2e-5 * G["image_features"].abs().max().item())
    assert torch.allclose(out["text_features"], G["text_features"], rtol=0, atol=2e-5 * G["text_features"].abs().max().item())
    assert abs(out["loss"].item() - G["loss"].item()) < 1e-5
    # per-block activations (forward hooks on the reference modules)
    for name, ref in G["acts"].items():
        tower, li = name[:3], int(name[3:])
        x = (out["vis_acts"] if tower == "vis" else out["txt_acts"])[li][:2, ::16, ::8]
        assert torch.allclose(x, ref, rtol=0, atol=1e-4 * ref.abs().max().item()), name
    # every gradient the reference's autograd produced
    assert set(out["grads"].keys()) == set(G["grads"].keys())
    # 145 at the default N_CTX=2 / PROMPT_DEPTH=9; 3 tensors per extra deep prompt (parameter + projection weight + bias)
    assert len(out["grads"]) == 145 + 3 * (m["depth"] - 9)
    worst = 0.0
    for n, packed in G["grads"].items():
        worst = max(worst, check_grad_against_golden(n, out["grads"][n], packed, rtol=2e-3))
    print("worst rel grad err", worst)


def test_fedavg_oracle_bit_exact_vs_reference():
    cases = load_golden("fedavg.pt")
    for K, c in cases.items():
        for key in c["inputs"][0]:
            mean32, mean16 = fedavg_oracle([d[key] for d in c["inputs"]])
            ref32 = c["fp32_mean"][key].reshape(-1)
            body = (ref32.numel() // 64) * 64  # vectorised body of torch's CPU cascade sum
            assert torch.equal(mean32.reshape(-1)[:body], ref32[:body]), (K, key)
            assert torch.allclose(mean32.reshape(-1), ref32, rtol=1e-5, atol=1e-6), (K, key)
            assert mean16.dtype == torch.float16
            assert torch.equal(mean16, c["out"][key]), (K, key)


@pytest.mark.parametrize("key", ["lr0.0026", "lr0.05"])
def test_oracle_three_step_trajectory_matches_reference(key):
    """The restated step (oracle forward/backward + clip_grad_norm_(1.0) + SGD momentum 0.9 / wd 5e-4, first-step
    momentum = gradient as in torch.optim.SGD) against the reference's own 3-step trajectory (tests/golden/
    traj_fp32.pt, trainers/maple.py:588-598): losses, pre-clip gradient norms and the logits after the updates."""
    torch.set_num_threads(8)
    G = load_golden("traj_fp32.pt")[key]
    m = G["meta"]
    sd, tok = customclip_state_dict(m["C"])
    orc = MapleOracle(sd, tok)
    mom, losses, norms = {}, [], []
    for s in range(m["steps"]):
        img, lab = synth.make_batch(m["B"], m["C"], m["seed_batch"] + s)
        out = orc.forward_backward(img, lab)
        g = out["grads"]
        norm = float(torch.sqrt(sum((v.double() ** 2).sum() for v in g.values())))
        coef = min(1.0, 1.0 / (norm + 1e-6))
        for k, v in g.items():
            d = v * coef + m["weight_decay"] * orc.P[k]
            mom[k] = d.clone() if k not in mom else mom[k].mul_(m["momentum"]).add_(d)
            orc.P[k].sub_(m["lr"] * mom[k])
        losses.append(out["loss"].item()); norms.append(norm)
    # fp32 re-association only; the larger LR amplifies it a little by the third step
    assert max(abs(a - b) / abs(b) for a, b in zip(losses, G["losses"])) < 1e-5
    assert max(abs(a - b) / abs(b) for a, b in zip(norms, G["grad_norms"])) < 3e-4
    img, _ = synth.make_batch(m["B"], m["C"], m["seed_batch"] + m["steps"])
    lg = orc.logits(img)
    assert (lg - G["logits_after"]).abs().max().item() < 1e-4 * G["logits_after"].abs().max().item()
    for name, packed in G["final"].items():
        if "full" in packed and name in orc.P:
            assert (orc.P[name] - packed["full"]).abs().max().item() <= 1e-4 * max(packed["full"].abs().max().item(), 1e-6), name
