#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 --launch-timeout 0 python scripts/race_k1q.py > gpurun_out/r2_sanitize_racecheck.log 2>&1
tail -40 gpurun_out/r2_sanitize_racecheck.log
