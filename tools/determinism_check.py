"""Run the same training steps twice from identical state and compare gradients / parameters bit for bit."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

dev = torch.device("cuda", 0); torch.cuda.set_device(0)
GRAPH = bool(int(os.environ.get("DC_GRAPH", "0")))
STEPS = int(os.environ.get("DC_STEPS", "4"))

def run():
    torch.manual_seed(1234); torch.cuda.manual_seed_all(1234)
    t = bench.make_trainer(dev, graph=GRAPH)
    eng = t.model.engine
    b = bench.host_batches(2, bench.B_PER_GPU, 0)
    snaps = []
    for s in range(STEPS):
        img, lab = b[s % 2]["img"].to(dev), b[s % 2]["label"].to(dev)
        t.step_async(img, lab)
        torch.cuda.synchronize()
        snaps.append((eng.grads.clone(), eng.params.clone(), t.read_step_result()[0]))
    return eng, snaps

e1, s1 = run()
e2, s2 = run()
for i, ((g1, p1, l1), (g2, p2, l2)) in enumerate(zip(s1, s2)):
    ge, pe = torch.equal(g1, g2), torch.equal(p1, p2)
    print(f"step {i}: loss {l1:.7f} / {l2:.7f}  grads equal {ge}  params equal {pe}")
    if not ge:
        for k, (off, n) in e1.offsets.items():
            a, b_ = g1[off:off + n], g2[off:off + n]
            if not torch.equal(a, b_):
                d = (a - b_).abs().max().item()
                print(f"   differs: {k}  max|d| {d:.3e}  (max|g| {a.abs().max().item():.3e})")
        break
