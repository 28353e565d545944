#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 1200 python -m pytest tests -q -m gpu --timeout 900 -s -x > gpurun_out/r02_tests.log 2>&1
echo "exit $?" >> gpurun_out/r02_tests.log
grep -E "passed|failed|error|exit" gpurun_out/r02_tests.log | tail -5
timeout 900 python bench.py --steps 50 --warmup 5 --no-c4 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
echo "bench rc $?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ["value","ms_per_step","e2e","gpu_launches"]}, d["roofline"]["frac"], d["roofline"]["step_frac"])
PY
