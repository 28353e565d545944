#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_trainers_gpu.py -q -m gpu --timeout 500 -x -k "fedavg or federated or check_finite" 2>&1 | tail -3
timeout 900 python bench.py --steps 50 --warmup 5 --no-c4 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
echo "bench rc $?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ["value","ms_per_step","fedavg_exchange_ms"]}, d["fedavg_roofline"])
PY
