#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "k1_instantiations or short_rotation or functional_bootstrap_dropin" 2>&1 | tail -15 | tee gpurun_out/pytest_k1c.log
timeout 600 python scripts/latency.py 2>&1 | tee gpurun_out/latency_k1c.log
