#!/usr/bin/env bash
# round 2, call A: k1q correctness + timing against k1 + one ncu capture of k1q at Level 1
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -x -q -k "k1_instantiations" 2>&1 | tail -15 | tee gpurun_out/r2a_pytest_k1q.log
timeout 600 python scripts/k1q_time.py 2>&1 | tee gpurun_out/r2a_k1q_time.log
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/r2a_pytest_gpu.log
WLS=level1 POLICIES=5 BATCH=4096 timeout 900 ncu --set full --clock-control none --import-source on -k regex:blind_rotate_k1q -s 1 -c 1 -f \
    -o gpurun_out/r2a_level1_k1q python scripts/k1q_time.py > gpurun_out/r2a_ncu_l1.log 2>&1
WLS=level2 POLICIES=5 BATCH=4096 timeout 900 ncu --set full --clock-control none --import-source on -k regex:blind_rotate_k1q -s 1 -c 1 -f \
    -o gpurun_out/r2a_level2_k1q python scripts/k1q_time.py > gpurun_out/r2a_ncu_l2.log 2>&1
ls -la gpurun_out
