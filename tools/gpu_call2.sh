#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 1200 python -m pytest tests -q -m gpu --timeout 900 -s > gpurun_out/r02_tests.log 2>&1
echo "exit $?" >> gpurun_out/r02_tests.log
grep -E "passed|failed|error|exit" gpurun_out/r02_tests.log | tail -5
python tools/exp_prec.py 2>&1 | grep -v Client | grep -A4 "0.0026 after"
