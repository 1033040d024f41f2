#!/usr/bin/env bash
# round 2, call G: k1q variant bits (1 = default, 3 = hybrid two-row pass B), handle-path timeline after the two-stream change
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "k1_instantiations or fullsize" 2>&1 | tail -5 | tee gpurun_out/r2g_pytest.log
for var in 1 3 2 0; do
  echo "== MB200_K1Q_VAR=$var"
  MB200_K1Q_VAR=$var POLICIES=5 timeout 600 python scripts/k1q_time.py 2>&1 | grep -v "fp64 peak"
done | tee gpurun_out/r2g_k1q_variants.log
# correctness of the hybrid variant on the instantiation test (forced through the env knob at the benchmark shapes only)
MB200_K1Q_VAR=3 timeout 900 python -m pytest tests -m gpu -x -q -k "fullsize or full_size_round_trip" 2>&1 | tail -3 | tee -a gpurun_out/r2g_pytest.log
timeout 600 python scripts/handle_trace.py 2> gpurun_out/r2g_handle_trace.log; grep -v "^$" gpurun_out/r2g_handle_trace.log | tail -24
