#!/usr/bin/env bash
# round 2, 8 GPUs: the multi-GPU mode inside the C library (drop-in C program, batch 32768) and the torchrun bench with
# the strong-scaling extra
set -x
NG=${NG:-8}
BATCH=${BATCH:-32768}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv | head -9
gcc -O1 -w -I include tests/dropin/dropin_multi.c -o /tmp/dropin_multi -L oracle/_ref -l:libmosfhet_avx512.so mosfhet_b200/libmosfhet_b200.so \
    -Wl,-rpath,$PWD/oracle/_ref -Wl,-rpath,$PWD/mosfhet_b200 -ldl -lm
LD_PRELOAD=$PWD/mosfhet_b200/libmosfhet_b200.so timeout 900 /tmp/dropin_multi $NG $BATCH 2>&1 | tail -3 | tee gpurun_out/r2_dropin_multi_${NG}gpu.log
MB200_TRACE=1 LD_PRELOAD=$PWD/mosfhet_b200/libmosfhet_b200.so timeout 900 /tmp/dropin_multi $NG $BATCH > /dev/null 2> gpurun_out/r2_dropin_multi_${NG}gpu_trace.log
[ -n "$SKIP_BENCH" ] && exit 0
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NG --steps 5 --warmup 3 2>gpurun_out/r2_bench_${NG}gpu.err | tail -1 | tee gpurun_out/r2_bench_level1_${NG}gpu.json
tail -3 gpurun_out/r2_bench_${NG}gpu.err
