"""Two eager training steps (BASELINE config 2: B=32, C=10) between cudaProfilerStart/Stop, for the ncu launch list:

    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
        --log-file gpurun_out/launches.csv python tools/profile_step.py
    python tools/launch_summary.py gpurun_out/launches.csv 2 > profiles/rNN_launches_eager_step_vXX_summary.txt

Eager (not graph-replayed) so that ncu sees every launch; per-launch times are cold-cache and serialised, so the
summary reports SHARES of the step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
t = bench.make_trainer(dev, graph=False)
b = bench.host_batches(1, bench.B_PER_GPU, 0)[0]
img, lab = b["img"].to(dev), b["label"].to(dev)
for _ in range(3):
    t.step_async(img, lab)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(2):
    t.step_async(img, lab)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profile_step done, loss", t.read_step_result()[0])
