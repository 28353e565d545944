"""Time of one training step in the fp32 parity mode vs the bf16 path (eager launches, B = 32, C = 10)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from helpers import customclip_state_dict
from federated_multi_modal_b200 import synth
from federated_multi_modal_b200.engine import MapleEngine

sd, tok = customclip_state_dict(10)
eng = MapleEngine(sd, tok)
img, lab = synth.make_batch(32, 10, 1)
img, lab = img.cuda(), lab.cuda()
for prec in ("bf16", "fp32"):
    for _ in range(3):
        eng.forward_backward(img, lab, precision=prec)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        eng.forward_backward(img, lab, precision=prec)
    e.record(); torch.cuda.synchronize()
    print(f"{prec}: {s.elapsed_time(e) / 10:.2f} ms per eager forward+backward at B=32, C=10")
