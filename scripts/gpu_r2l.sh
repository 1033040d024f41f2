#!/usr/bin/env bash
# round 2, call L: N = 2048 back on the slot mapping; VAR = 3 against 7 at both levels (segments on)
set -x
mkdir -p gpurun_out
for var in 3 7; do
  echo "== MB200_K1Q_VAR=$var"
  MB200_K1Q_VAR=$var POLICIES=5 timeout 600 python scripts/k1q_time.py 2>&1 | grep -v "fp64 peak"
done | tee gpurun_out/r2l_k1q_variants.log
timeout 900 python -m pytest tests -m gpu -x -q -k "k1_instantiations or fullsize or full_size_round_trip or segmented" 2>&1 | tail -3 | tee gpurun_out/r2l_pytest.log
