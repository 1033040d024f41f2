#!/usr/bin/env bash
# Runs on the GPU box: circuit bootstrap / table key switch tests only (fast iteration).
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "vertical_packing" 2>&1 | tail -40 | tee gpurun_out/pytest_cb.log
