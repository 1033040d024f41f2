#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 900 python scripts/k1h_variants.py 2>&1 | tee gpurun_out/k1h_variants.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
