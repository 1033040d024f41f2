#!/usr/bin/env bash
# round 2, call N: Level 2 with the key segment pinned in L2 (access policy window), against the plain segmented run
set -x
mkdir -p gpurun_out
for mode in plain persist; do
  for mb in 56 42 28; do
    echo "== $mode budget $mb MB"
    if [ $mode = persist ]; then export MB200_L2_PERSIST=1; else unset MB200_L2_PERSIST; fi
    MB200_SEG_BUDGET_MB=$mb WLS=level2 POLICIES=5 timeout 600 python scripts/k1q_time.py 2>&1 | grep -v "fp64 peak"
  done
done | tee gpurun_out/r2n_l2_persist.log
