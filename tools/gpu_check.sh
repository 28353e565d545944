#!/bin/bash
# Runs on the GPU box via gpurun: tests under a timeout so a hung kernel cannot eat the box.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
T=${1:-tests}; shift
timeout 1500 python -m pytest $T -x -q -m gpu --timeout 600 -s "$@" > gpurun_out/tests.log 2>&1
echo "exit $?" >> gpurun_out/tests.log
tail -60 gpurun_out/tests.log
