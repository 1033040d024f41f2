#!/usr/bin/env bash
# round 2, call O: flat host entry point pipelined like the handle path; smoke with both kernels; Level-2 DRAM bytes against the segment budget
set -x
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2 | tee gpurun_out/r2o_smoke.log
timeout 900 python -m pytest tests -m gpu -x -q -k "round_trip or host or ragged or fullsize or per_input" 2>&1 | tail -3 | tee gpurun_out/r2o_pytest.log
timeout 900 python bench.py --no-extras --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2o_bench_level1.json
timeout 900 python bench.py --no-extras --no-cpu-baseline --workload level2 --steps 3 2>/dev/null | tail -1 > gpurun_out/r2o_bench_level2.json
python - <<'PY'
import json
for f in ("level1", "level2"):
    d = json.load(open(f"gpurun_out/r2o_bench_{f}.json"))
    print(f, "value", d["value"], "e2e", d["e2e"]["value"], "ratio", d["e2e"]["value"] / d["value"], "match", d["e2e"].get("matches_device_path"))
PY
for mb in 56 42 28; do
  MB200_SEG_BUDGET_MB=$mb WLS=level2 POLICIES=5 BATCH=4096 timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
      --clock-control none -k regex:blind_rotate_k1q -c 30 --csv --log-file gpurun_out/r2o_level2_dram_budget${mb}.csv python scripts/k1q_time.py > /dev/null 2>&1
done
