#!/usr/bin/env bash
# round 2, call F (2 GPUs): multi-GPU mode inside the library, torchrun bench with the strong-scaling extra, handle-path timeline
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
timeout 600 python scripts/handle_trace.py 2> gpurun_out/r2f_handle_trace.log; tail -40 gpurun_out/r2f_handle_trace.log
timeout 900 python -m pytest tests/test_dropin_preload.py -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/r2f_pytest_dropin.log
# the multi-GPU drop-in program at the benchmark parameters, batch 8192 over 2 devices
gcc -O1 -w -I include tests/dropin/dropin_multi.c -o /tmp/dropin_multi -L oracle/_ref -l:libmosfhet_avx512.so mosfhet_b200/libmosfhet_b200.so \
    -Wl,-rpath,$PWD/oracle/_ref -Wl,-rpath,$PWD/mosfhet_b200 -ldl -lm
LD_PRELOAD=$PWD/mosfhet_b200/libmosfhet_b200.so timeout 900 /tmp/dropin_multi 2 8192 2>&1 | tail -3 | tee gpurun_out/r2f_dropin_multi_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 2>gpurun_out/r2f_bench_2gpu.err | tail -1 | tee gpurun_out/r2f_bench_2gpu.json
tail -3 gpurun_out/r2f_bench_2gpu.err
