#!/usr/bin/env bash
# round 2, call D: k1q v4 (digit parking at N = 2048), extraction family, stale keys, bench
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "k1_instantiations" 2>&1 | tail -5 | tee gpurun_out/r2d_pytest_k1q.log
for kpf in 1 0; do
  echo "== MB200_K1Q_KPF=$kpf"
  MB200_K1Q_KPF=$kpf timeout 600 python scripts/k1q_time.py 2>&1
done | tee gpurun_out/r2d_k1q_time.log
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 | tee gpurun_out/r2d_pytest_gpu.log
WLS=level1 POLICIES=5 BATCH=4096 timeout 900 ncu --set full --clock-control none --import-source on -k regex:blind_rotate_k1q -s 1 -c 1 -f \
    -o gpurun_out/r2d_level1_k1q python scripts/k1q_time.py > gpurun_out/r2d_ncu_l1.log 2>&1
WLS=level2 POLICIES=5 BATCH=4096 timeout 900 ncu --set full --clock-control none --import-source on -k regex:blind_rotate_k1q -s 1 -c 1 -f \
    -o gpurun_out/r2d_level2_k1q python scripts/k1q_time.py > gpurun_out/r2d_ncu_l2.log 2>&1

timeout 600 python bench.py --steps 5 --warmup 3 --cpu-seconds 6 2>&1 | tail -1 | tee gpurun_out/r2d_bench_level1.json
timeout 600 python bench.py --steps 3 --warmup 3 --workload level2 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r2d_bench_level2.json
ls -la gpurun_out
