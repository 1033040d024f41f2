#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
timeout 900 python scripts/k1_variants.py 2>&1 | tee gpurun_out/k1_variants.log
