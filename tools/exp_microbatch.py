"""Experiment: one batch of 32 as a single pass vs two concurrent half-batches on two streams (fills launch bubbles?)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from helpers import customclip_state_dict
from federated_multi_modal_b200 import synth
from federated_multi_modal_b200.engine import MapleEngine

dev = torch.device("cuda", 0); torch.cuda.set_device(0)
sd, tok = customclip_state_dict(10)
A = MapleEngine(sd, tok)
Bm = MapleEngine(sd, tok, share_from=A, share_workspace=False)
img, lab = synth.make_batch(32, 10, 1)
img, lab = img.cuda(), lab.cuda()
s2 = torch.cuda.Stream()

def full():
    A.forward_backward(img, lab)
def halves():
    main = torch.cuda.current_stream()
    s2.wait_stream(main)
    with torch.cuda.stream(s2):
        Bm.forward_backward(img[16:], lab[16:])
    A.forward_backward(img[:16], lab[:16])
    main.wait_stream(s2)

def bench(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20): g.replay()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / 20
print("single pass, B=32      : %.3f ms" % bench(full))
print("two concurrent halves  : %.3f ms" % bench(halves))
