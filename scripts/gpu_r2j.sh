#!/usr/bin/env bash
# round 2, call J: segment budget sweep (key bytes per launch) at both levels
set -x
mkdir -p gpurun_out
for mb in 400 88 60 44 30 22 16; do
  echo "== MB200_SEG_BUDGET_MB=$mb"
  MB200_SEG_BUDGET_MB=$mb POLICIES=5 timeout 600 python scripts/k1q_time.py 2>&1 | grep -v "fp64 peak"
done | tee gpurun_out/r2j_segment_sweep.log
