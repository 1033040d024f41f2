#!/usr/bin/env bash
# round 2, call P: key switch sliced only where it pays; 4 key segments at Level 2; e2e at both levels; handle path at Level 2
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "round_trip or host or ragged or fullsize or per_input or segmented or handle or dropin" 2>&1 | tail -3 | tee gpurun_out/r2p_pytest.log
timeout 900 python bench.py --no-extras --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2p_bench_level1.json
timeout 900 python bench.py --no-extras --no-cpu-baseline --workload level2 --steps 3 2>/dev/null | tail -1 > gpurun_out/r2p_bench_level2.json
python - <<'PY'
import json
for f in ("level1", "level2"):
    d = json.load(open(f"gpurun_out/r2p_bench_{f}.json"))
    print(f, "value", d["value"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "ratio", d["e2e"]["value"] / d["value"], "match", d["e2e"].get("matches_device_path"))
PY
WL=level2 BATCH=4096 timeout 600 python scripts/handle_trace.py 2>&1 | tail -9
