#!/usr/bin/env bash
# round 2: compute-sanitizer over the new code -- memcheck on the k1q instantiations, the extraction family, the fused digit
# step and the stale-key path; racecheck (shared-memory hazards; the kernel leans on warp-level barriers) on the k1q kernels
set -x
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 0 \
  python -m pytest tests/test_gpu_parity.py -x -q -m gpu \
  -k "k1_instantiations_vs_generic or mv_extract or integer_digit_step or stale_key or programmable_bootstrap_with_unfolded" 2>&1 | tail -12 | tee gpurun_out/r2_sanitize_memcheck.log
# racecheck: k1q at N = 1024 (l = 3: 2 + 1 level batches, hybrid two-row phases) and N = 2048 (l = 4), short rotations
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 --launch-timeout 0 \
  python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "k1_instantiations_vs_generic and (1024 or 2048)" 2>&1 | tail -12 | tee gpurun_out/r2_sanitize_racecheck.log
