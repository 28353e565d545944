"""Graph-replayed time of each tower's forward and backward alone (is the side-stream text tower ever the longer branch?)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from helpers import customclip_state_dict
from federated_multi_modal_b200 import synth
from federated_multi_modal_b200.engine import MapleEngine

dev = torch.device("cuda", 0); torch.cuda.set_device(0)
sd, tok = customclip_state_dict(10)
eng = MapleEngine(sd, tok)
img, lab = synth.make_batch(32, 10, 1)
img, lab = img.cuda(), lab.cuda()
eng.forward_backward(img, lab); torch.cuda.synchronize()
dfi, dft = eng.last["dfi"].clone(), eng.last["dft"].clone()

def tfwd():
    eng._prompt_learner_fwd(); return eng._text_features(True)
def vfwd():
    return eng._image_features(img, True)
state = {}
def tbwd():
    for tw in (eng.vis, eng.txt): tw.ln_slot, tw.ln_pending = 0, []
    ft, txs, tstat = state["t"]
    eng._tower_bwd(eng.txt, dft, eng.tproj, txs, tstat, eng.txt_rows, "text_encoder.ln_final", eng.C, 1)
    eng._ln_reduce(eng.txt)
def vbwd():
    for tw in (eng.vis, eng.txt): tw.ln_slot, tw.ln_pending = 0, []
    fi, vxs, vstat = state["v"]
    eng._tower_bwd(eng.vis, dfi, eng.vproj, vxs, vstat, eng.cls_rows, "image_encoder.ln_post", 32, eng.Tv - eng.n)
    eng._ln_reduce(eng.vis)

def bench(fn, name):
    for _ in range(2): fn()
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
    print(f"{name:18s} {s.elapsed_time(e) / 20:7.3f} ms")

with torch.no_grad():
    state["t"] = tfwd(); state["v"] = vfwd()
    bench(tfwd, "text forward"); bench(vfwd, "vision forward")
    state["t"] = tfwd(); state["v"] = vfwd()
    bench(tbwd, "text backward"); bench(vbwd, "vision backward")
