#!/usr/bin/env bash
# memcheck over the full-size composed paths, racecheck (shared-memory hazards) over the k = 1 kernels
set -x
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 0 \
  python -m pytest tests/test_gpu_parity.py -x -q -m gpu \
  -k "circuit_bootstrap_full_size or trgsw_accumulator_bootstrap_full_size or unfolded_bootstrap_full_size or cmux_and_vertical" 2>&1 | tail -8 | tee gpurun_out/sanitize_full.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 --launch-timeout 0 \
  python -m pytest tests/test_gpu_parity.py -x -q -m gpu \
  -k "cmux_and_vertical or short_rotation or test_keyswitch_bit_exact" 2>&1 | tail -8 | tee gpurun_out/sanitize_race.log
