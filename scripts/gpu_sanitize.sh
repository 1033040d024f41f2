#!/usr/bin/env bash
# compute-sanitizer memcheck over the small-parameter parity tests (fixtures with N <= 256): out-of-bounds and
# misaligned accesses in every kernel family, including the composed paths of SURVEY 8(f)
set -x
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 0 \
  python -m pytest tests/test_gpu_parity.py -x -q -m gpu \
  -k "dropin or bit_exact or transforms or extprod_flat or edge_cases or multivalue or batched" 2>&1 | tail -25 | tee gpurun_out/sanitize.log
