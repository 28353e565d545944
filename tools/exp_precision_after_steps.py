import sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from test_parity_benched_gpu import _trainer
from federated_multi_modal_b200 import synth
import torch.nn.functional as Fn
def ev(eng, img):
    a = eng.logits(img, cache_text=False).clone(); fa = eng.last_image_features()
    fta = eng._ft_cache.clone()
    b = eng.logits(img, precision="fp32").clone(); fb = eng.last_image_features()
    na, nb = Fn.normalize(fa, dim=-1), Fn.normalize(fb, dim=-1)
    t = Fn.normalize(fta, dim=-1)
    d = na - nb
    return (a - b).abs().max().item(), b.abs().max().item(), d.norm(dim=1).max().item(), ((d @ t.t()).abs().max() / d.norm(dim=1).max()).item(), (t @ t.t()).min().item()
for lr in (0.0026, 0.05):
    t = _trainer(10, graph=False, lr=lr)
    eng = t.model.engine
    print("lr", lr, "init:")
    for seed in (503, 123, 77, 9):
        img = synth.make_batch(4, 10, seed)[0].cuda()
        print("   seed", seed, "abs err %.4f max|l| %.3f |d_img_n| %.2e  align %.3f  min cos(t,t) %.3f" % ev(eng, img))
    for s in range(3):
        img, lab = synth.make_batch(4, 10, 500 + s)
        t.optim.lr = lr
        t.forward_backward({"img": img, "label": lab})
    print("lr", lr, "after 3 steps:")
    for seed in (503, 123, 77, 9):
        img = synth.make_batch(4, 10, seed)[0].cuda()
        print("   seed", seed, "abs err %.4f max|l| %.3f |d_img_n| %.2e  align %.3f  min cos(t,t) %.3f" % ev(eng, img))
