#!/usr/bin/env bash
# final scaling lines of the round on an 8-GPU box: level1 at N = 8, 4, 2 and config[2] (level2, 65536 ciphertexts) at N = 8
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for N in 8 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N \
   bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_level1_${N}gpu.json
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 \
   bench.py --gpus 8 --steps 2 --warmup 3 --workload level2 --batch 8192 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_level2_8gpu_b8192.json
