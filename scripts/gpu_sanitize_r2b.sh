#!/usr/bin/env bash
# round 2, end: compute-sanitizer over what changed after the first pass -- key switch sliced by blockIdx.y (atomics into a
# zeroed output), the pipelined flat and handle paths (two streams, sliced key switch, overlapped copies), key segments,
# k1q with folded butterflies (racecheck)
set -x
mkdir -p gpurun_out
timeout 2000 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 0 \
  python -m pytest tests -x -q -m gpu \
  -k "keyswitch_bit_exact or pipelined_host or table_keyswitches or segmented or full_size_round_trip" 2>&1 | tail -12 | tee gpurun_out/r2b_sanitize_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 --launch-timeout 0 python scripts/race_k1q.py > gpurun_out/r2b_sanitize_racecheck.log 2>&1
tail -12 gpurun_out/r2b_sanitize_racecheck.log
