#!/bin/bash
mkdir -p gpurun_out
python tools/attn_bench.py 2>&1 | tail -2
for i in 1 2; do
timeout 900 python bench.py --steps 100 --warmup 10 --no-c4 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ["value","ms_per_step"]})
PY
done
