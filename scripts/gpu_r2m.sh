#!/usr/bin/env bash
# round 2, call M: handle-array path in prioritised groups (bootstrap -> key switch -> copy back per group)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "handle or dropin or fullsize or keyswitch or stale or multi or circuit" 2>&1 | tail -3 | tee gpurun_out/r2m_pytest.log
timeout 600 python scripts/ks_sweep.py 2>&1 | tee gpurun_out/r2m_ks_sweep.log
timeout 600 python scripts/handle_trace.py 2> gpurun_out/r2m_handle_trace.log
tail -22 gpurun_out/r2m_handle_trace.log
timeout 900 python bench.py 2>gpurun_out/r2m_bench.err | tail -1 > gpurun_out/r2m_bench_level1.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2m_bench_level1.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "handles", d.get("e2e_handles"), "parity", str(d.get("parity"))[:300])
PY
