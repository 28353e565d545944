#!/bin/bash
# A/B of the graph-replayed training step between library builds (same box, same session): args = lib paths
for lib in "$@"; do
  echo "== $lib"
  MFK_LIB_PATH=$lib python - <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
import bench
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
t = bench.make_trainer(dev, graph=True)
b = bench.host_batches(1, 32, 0)[0]
img, lab = b["img"].to(dev), b["label"].to(dev)
for _ in range(8): t.step_async(img, lab)
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(30): t.step_async(img, lab)
    e.record(); torch.cuda.synchronize()
    best = min(best, s.elapsed_time(e) / 30)
print("ms/step %.3f  loss %.5f" % (best, t.read_step_result()[0]))
PY
done
