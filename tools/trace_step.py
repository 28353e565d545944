"""Kernel-level timeline of graph-replayed training steps via torch.profiler (CUPTI): warm, in-situ durations."""
import os, sys, collections, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
t = bench.make_trainer(dev, graph=True)
b = bench.host_batches(1, bench.B_PER_GPU, 0)[0]
img, lab = b["img"].to(dev), b["label"].to(dev)
for _ in range(5):
    t.step_async(img, lab)
torch.cuda.synchronize()
N = 5
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(N):
        t.step_async(img, lab)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
agg = collections.defaultdict(lambda: [0, 0.0])
tmin, tmax = 1e30, 0
for e in ev:
    name = e.name.replace("void ", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    name = name.split("(")[0]
    agg[name][0] += 1
    agg[name][1] += e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total
    tr = e.time_range
    tmin, tmax = min(tmin, tr.start), max(tmax, tr.end)
# per-stream occupancy: union of kernel intervals vs the step span (idle = launch gaps / dependencies)
by_stream = collections.defaultdict(list)
for e in ev:
    by_stream[getattr(e, "device_index", 0), getattr(e, "stream", None) or getattr(e, "device_resource_id", 0)].append(
        (e.time_range.start, e.time_range.end))
for k, iv in sorted(by_stream.items(), key=lambda kv: -len(kv[1])):
    iv.sort()
    busy, cur_s, cur_e = 0.0, iv[0][0], iv[0][1]
    for a, b_ in iv[1:]:
        if a > cur_e:
            busy += cur_e - cur_s
            cur_s, cur_e = a, b_
        else:
            cur_e = max(cur_e, b_)
    busy += cur_e - cur_s
    print(f"stream {k}: {len(iv) / N:.0f} kernels/step, busy {busy / N:.1f} us/step, summed durations "
          f"{sum(b_ - a for a, b_ in iv) / N:.1f} us/step")
tot = sum(v[1] for v in agg.values())
print(f"steps {N}: span {(tmax - tmin) / N:.1f} us/step, sum of kernel time {tot / N:.1f} us/step, kernels/step {len(ev) / N:.0f}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{v[1] / N:9.1f} us/step {v[0] / N:6.1f} x {v[1] / v[0]:7.2f} us  {k[:80]}")
