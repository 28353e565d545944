"""Step time / throughput / roofline fraction of the engine (fwd+bwd+clip+SGD) versus per-GPU batch size."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from helpers import customclip_state_dict
from federated_multi_modal_b200 import synth
from federated_multi_modal_b200.engine import MapleEngine

dev = torch.device("cuda", 0); torch.cuda.set_device(0)
sd, tok = customclip_state_dict(10)
eng = MapleEngine(sd, tok)
peak = 1392.1e12
out = []
for B in [int(x) for x in (sys.argv[1:] or ["16", "32", "64", "128", "256"])]:
    img, lab = synth.make_batch(B, 10, 1)
    img, lab = img.cuda(), lab.cuda()
    hyper = torch.tensor([0.0026, 0.9, 0.0, 5e-4, 1.0, 0.0, 0.0], device=dev)
    def step():
        eng.forward_backward(img, lab)
        eng.sgd_step(0.0, hyper=hyper)
    for _ in range(2): step()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        step()
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10): g.replay()
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 10
    fl = eng.flops_per_step(B)
    out.append(dict(B=B, ms=ms, img_s=B / ms * 1e3, frac=fl / (ms * 1e-3) / peak))
    print(f"B={B:4d}  {ms:8.3f} ms/step  {B / ms * 1e3:9.1f} img/s  {fl / (ms * 1e-3) / 1e12:7.1f} TFLOP/s  ({100 * fl / (ms * 1e-3) / peak:4.1f}% of sustained peak)")
json.dump(out, open("gpurun_out/batch_sweep.json", "w"))
