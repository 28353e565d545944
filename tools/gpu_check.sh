#!/bin/bash
# Runs on the GPU box via gpurun: kernel tests (each under a timeout so a hung kernel cannot eat the box).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu --timeout 300 "$@" > gpurun_out/kernels.log 2>&1
echo "exit $?" >> gpurun_out/kernels.log
tail -40 gpurun_out/kernels.log
