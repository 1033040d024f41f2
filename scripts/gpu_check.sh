#!/usr/bin/env bash
# Runs on the GPU box (via gpurun): parity tests, smoke, short benches.  Outputs land in gpurun_out/.
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -5
timeout 600 python bench.py --steps 3 --warmup 3 ${BENCH_ARGS:---cpu-seconds 6} 2>&1 | tail -2 | tee gpurun_out/bench_level1.json
timeout 600 python bench.py --steps 2 --warmup 3 --workload level2 --no-cpu-baseline 2>&1 | tail -2 | tee gpurun_out/bench_level2.json
timeout 600 python scripts/latency.py 2>&1 | tee gpurun_out/latency.log
