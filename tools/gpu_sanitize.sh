#!/bin/bash
# compute-sanitizer over the kernel tests: memcheck on the whole kernel suite, racecheck + synccheck on the
# kernels with cross-warp protocols (split-K GEMM tails, tcgen05 attention forward / fused backward, LayerNorm).
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
run() { # name tool timeout pytest-args...
  local name=$1 tool=$2 to=$3; shift 3
  timeout $to $CS --tool $tool --launch-timeout 0 --print-limit 50 --error-exitcode 86 \
    python -m pytest "$@" -q -x --timeout 3000 -p no:cacheprovider > gpurun_out/r02_sanitizer_$name.log 2>&1
  echo "$name rc $?" | tee -a gpurun_out/r02_sanitizer_$name.log
  grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY|Hazard|Invalid|rc " gpurun_out/r02_sanitizer_$name.log | tail -6
}
run memcheck memcheck 700 tests/test_kernels_gpu.py -m gpu
run racecheck racecheck 500 tests/test_kernels_gpu.py -m gpu -k "splitk_tail or (attention_fwd_bwd and 199) or layernorm_fwd_bwd or test_head"
