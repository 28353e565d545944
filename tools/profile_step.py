"""One or two eager (non-graph) training steps of the bench workload, for ncu launch lists / captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
t = bench.make_trainer(dev, graph=False)
b = bench.host_batches(1, bench.B_PER_GPU, 0)[0]
img, lab = b["img"].to(dev), b["label"].to(dev)
for _ in range(steps):
    t.step_async(img, lab)
torch.cuda.synchronize()
print("loss", t.read_step_result()[0])
