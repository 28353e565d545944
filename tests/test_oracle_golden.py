"""CPU: pins oracle/maple_cpu.py to the golden vectors generated from the unmodified reference."""
import pytest
import torch

from oracle.maple_cpu import MapleOracle, fedavg_oracle
from federated_multi_modal_b200 import synth
from helpers import load_golden, customclip_state_dict, check_grad_against_golden


@pytest.mark.parametrize("fixture", ["c1_fp32.pt", "c3s_fp32.pt"])
def test_oracle_matches_reference_autograd(fixture):
    torch.set_num_threads(8)
    G = load_golden(fixture)
    m = G["meta"]
    sd, tok = customclip_state_dict(m["C"], m["seed_clip"], m["seed_pl"])
    orc = MapleOracle(sd, tok)
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
    assert len(out["grads"]) == 145
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
