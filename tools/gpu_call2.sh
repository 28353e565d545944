#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 -s > gpurun_out/r02_tests.log 2>&1
echo "exit $?" >> gpurun_out/r02_tests.log
grep -E "passed|failed|error|exit" gpurun_out/r02_tests.log | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
echo "bench rc $?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ["value","ms_per_step","e2e","gpu_launches","fedavg_exchange_ms","fedavg_round_s"]}, d["roofline"]["frac"], d["roofline"]["step_frac"], d["fed_round_c4"])
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.json 2>/dev/null; tail -c 600 gpurun_out/r02_bench_ref.json
DC_GRAPH=1 DC_STEPS=6 python tools/determinism_check.py 2>&1 | tail -6 > gpurun_out/r02_determinism_graph.log; tail -2 gpurun_out/r02_determinism_graph.log
