#!/usr/bin/env bash
# round 2, call Q: N = 2048 A' split over the two slots of a polynomial by output parity (all warps busy, no exchange)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "k1_instantiations or fullsize or full_size_round_trip or segmented" 2>&1 | tail -3 | tee gpurun_out/r2q_pytest.log
for var in 3 7; do
  echo "== MB200_K1Q_VAR=$var"
  MB200_K1Q_VAR=$var WLS=level2 POLICIES=5 timeout 600 python scripts/k1q_time.py 2>&1 | grep -v "fp64 peak"
done | tee gpurun_out/r2q_k1q_variants.log
MB200_NO_SEGMENTS=1 WLS=level2 POLICIES=5 BATCH=4096 timeout 900 ncu --set full --clock-control none --import-source on -k regex:blind_rotate_k1q -s 1 -c 1 -f \
    -o gpurun_out/r2q_level2_k1q python scripts/k1q_time.py > gpurun_out/r2q_ncu_l2.log 2>&1
