#!/bin/bash
# round 2, call 1: baseline GPU tests + ncu --set full of every non-GEMM kernel class on real operands
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu --timeout 600 > gpurun_out/r02_tests_base.log 2>&1
echo "exit $?" >> gpurun_out/r02_tests_base.log
tail -3 gpurun_out/r02_tests_base.log
timeout 300 python tools/ncu_targets.py > gpurun_out/ncu_targets_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:'attn_|ln_|vis_assemble|head_|sumsq|sgd_step|fedavg|check_finite|gemm_bf16' \
  -o gpurun_out/r02_targets_v16 -f python tools/ncu_targets.py > gpurun_out/ncu_targets.log 2>&1
echo "ncu rc $?"
tail -5 gpurun_out/ncu_targets.log
ls -la gpurun_out/*.ncu-rep
