#!/bin/bash
# ncu evidence for the current build: (1) launch list of two eager steps, (2) --set full of one launch of every kernel
# class on real operands. Each only after the same command has exited 0 without ncu.
V=${1:-v18}
mkdir -p gpurun_out
timeout 300 python tools/profile_step.py > gpurun_out/profile_step_plain.log 2>&1 || { echo "plain step failed"; tail -5 gpurun_out/profile_step_plain.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/r02_launches_eager_step_$V.csv python tools/profile_step.py > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc $?"; wc -l gpurun_out/r02_launches_eager_step_$V.csv
timeout 300 python tools/ncu_targets.py > gpurun_out/ncu_targets_plain.log 2>&1 || { echo "plain targets failed"; tail -5 gpurun_out/ncu_targets_plain.log; exit 1; }
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:'attn_|ln_|vis_assemble|head_|sumsq|sgd_step|fedavg|check_finite|gemm_bf16|scatter_rows' \
  -o gpurun_out/r02_targets_$V -f python tools/ncu_targets.py > gpurun_out/ncu_targets.log 2>&1
echo "ncu targets rc $?"; tail -3 gpurun_out/ncu_targets.log; ls -la gpurun_out/*.ncu-rep
