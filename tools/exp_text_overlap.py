import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from federated_multi_modal_b200 import engine as E

dev = torch.device("cuda", 0); torch.cuda.set_device(0)
def run(skip, only=None):
    t = bench.make_trainer(dev, graph=True)
    eng = t.model.engine
    if skip:
        orig_tf, orig_tb = eng._text_features, eng._tower_bwd
        cache = {}
        def tf(train, class_range=None):
            if "v" not in cache: cache["v"] = orig_tf(train, class_range)
            return cache["v"]
        def tb(tw, *a, **k):
            if tw is eng.txt:
                if "b" not in cache: cache["b"] = orig_tb(tw, *a, **k)
                return cache["b"]
            return orig_tb(tw, *a, **k)
        if only in (None, "fwd"): eng._text_features = tf
        if only in (None, "bwd"): eng._tower_bwd = tb
    b = bench.host_batches(1, 32, 0)[0]
    img, lab = b["img"].to(dev), b["label"].to(dev)
    for _ in range(5): t.step_async(img, lab)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20): t.step_async(img, lab)
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / 20
print("with text tower   : %.3f ms/step" % run(False))
print("text tower skipped: %.3f ms/step" % run(True))
print("text FORWARD skipped only : %.3f ms/step" % run(True, "fwd"))
print("text BACKWARD skipped only: %.3f ms/step" % run(True, "bwd"))
