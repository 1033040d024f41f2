#!/usr/bin/env bash
# ncu captures for both workloads, summarised on the box (the .ncu-rep files are too large to bring back together)
TAG=${1:-r1q}
mkdir -p gpurun_out
for WL in level1 level2; do
  bash scripts/gpu_profile.sh $TAG $WL > gpurun_out/prof_${TAG}_${WL}.log 2>&1
  for K in k1 ks; do
    python scripts/ncu_summary.py gpurun_out/${TAG}_${WL}_${K}.ncu-rep > gpurun_out/${TAG}_${WL}_${K}_ncu_summary.txt 2>&1
    rm -f gpurun_out/${TAG}_${WL}_${K}.ncu-rep
  done
done
timeout 600 python bench.py --steps 3 --warmup 3 --cpu-seconds 6 2>&1 | tail -1 > gpurun_out/bench_level1.json
timeout 600 python bench.py --steps 2 --warmup 3 --workload level2 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_level2.json
ls -la gpurun_out | grep $TAG
