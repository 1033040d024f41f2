#!/usr/bin/env bash
# round 2, call E: full GPU suite (extraction family, multi-GPU mode on one device, pipelined handle path) + the bench line with extras
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 | tee gpurun_out/r2e_pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 --cpu-seconds 6 2>gpurun_out/r2e_bench_level1.err | tail -1 | tee gpurun_out/r2e_bench_level1.json
tail -5 gpurun_out/r2e_bench_level1.err
timeout 600 python bench.py --steps 3 --warmup 3 --workload level2 --no-cpu-baseline --no-extras 2>&1 | tail -1 | tee gpurun_out/r2e_bench_level2.json
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
