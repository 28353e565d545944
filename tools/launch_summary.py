"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) of tools/profile_step.py per kernel / grid /
stream: launches per step, mean duration, share of the step.  python tools/launch_summary.py launches.csv [steps]"""
import csv, re, sys
from collections import defaultdict

path, steps = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 2
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ix = {h: i for i, h in enumerate(hdr)}
recs = [(r[ix["Kernel Name"]], r[ix["Grid Size"]], r[ix["Stream"]], float(r[ix["Metric Value"]]) / 1e3) for r in rows]
# keep the launches of the training steps only: everything from the first libmfk kernel of the first step on
mfk = [i for i, r in enumerate(recs) if "mfk" in r[0] or "gemm_bf16" in r[0] or "ln_" in r[0] or "attn_" in r[0]]
recs = recs[mfk[0]:] if mfk else recs
ours = recs  # library (ATen / memcpy) kernels inside the steps are reported too, tagged [lib]
n_lib = sum(1 for r in recs if r[0].startswith("void at::"))
per = len(ours) / steps
agg = defaultdict(lambda: [0, 0.0])
for name, grid, stream, us in ours:
    short = re.sub(r"^void |\(anonymous namespace\)::|mfk::|\(.*$", "", name)
    if name.startswith("void at::"):
        short = "[lib] " + short
    k = (short, grid, stream)
    agg[k][0] += 1
    agg[k][1] += us
total = sum(v[1] for v in agg.values()) / steps
print(f"# ncu launch list of {steps} eager training steps ({path}), aggregated per kernel / grid / stream")
print("# cold-cache, serialised: compare SHARES, not absolute times")
print(f"# {per:.0f} launches and {total:.1f} us per step (sum over both streams); {n_lib / steps:.1f} of them per step are "
      f"library (ATen) kernels, tagged [lib]\n")
for (name, grid, stream), (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name[:58]:58s} {grid:14s} s{stream:3s} {n / steps:6.1f}/step x {us / n:7.1f} us = {us / steps:8.1f} us/step {100 * us / steps / total:5.1f}%")
