#!/usr/bin/env bash
# $1 = number of GPUs
set -x
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_dropin_preload.py -m gpu -x -q 2>&1 | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps 3 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench_level1_${N}gpu.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
   bench.py --gpus $N --steps 2 --warmup 3 --workload level2 2>&1 | tail -3 | tee gpurun_out/bench_level2_${N}gpu.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
   bench.py --impl reference --gpus $N --steps 1 --warmup 1 --cpu-seconds 3 2>&1 | tail -2
