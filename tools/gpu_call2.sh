#!/bin/bash
# round 2, call 2: new parity tests + full GPU suite + bench line
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 1200 python -m pytest tests -x -q -m gpu --timeout 900 -s > gpurun_out/r02_tests.log 2>&1
echo "exit $?" >> gpurun_out/r02_tests.log
grep -E "passed|failed|error|exit" gpurun_out/r02_tests.log | tail -5
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
echo "bench rc $?"
tail -c 3000 gpurun_out/r02_bench.json
tail -5 gpurun_out/r02_bench.err
