#!/usr/bin/env bash
# round 2, call H: k1q variant bits with folded butterflies (4), on top of key pipelining (1) and the hybrid two-row pass B (2)
set -x
mkdir -p gpurun_out
for var in 3 7 5; do
  echo "== MB200_K1Q_VAR=$var"
  MB200_K1Q_VAR=$var POLICIES=5 timeout 600 python scripts/k1q_time.py 2>&1 | grep -v "fp64 peak"
done | tee gpurun_out/r2h_k1q_variants.log
MB200_K1Q_VAR=7 timeout 900 python -m pytest tests -m gpu -x -q -k "fullsize or full_size_round_trip" 2>&1 | tail -3 | tee gpurun_out/r2h_pytest.log
MB200_K1Q_VAR=7 WLS=level1 POLICIES=5 BATCH=4096 timeout 900 ncu --set full --clock-control none --import-source on -k regex:blind_rotate_k1q -s 1 -c 1 -f \
    -o gpurun_out/r2h_level1_k1q python scripts/k1q_time.py > gpurun_out/r2h_ncu_l1.log 2>&1
