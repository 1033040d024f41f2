#!/usr/bin/env bash
# round 2, final evidence (one GPU): full GPU test suite, smoke, bench lines (both arms, both levels), launch lists of the
# bench command, ncu --set full of the dominant kernel and the key switch at both levels.  Outputs: gpurun_out/r2z_*
set -x
T=r2z
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/${T}_pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3 | tee gpurun_out/${T}_smoke.log
timeout 900 python bench.py 2>gpurun_out/${T}_bench_level1.err | tail -1 > gpurun_out/${T}_bench_level1.json
timeout 900 python bench.py --impl reference 2>gpurun_out/${T}_bench_reference.err | tail -1 > gpurun_out/${T}_bench_reference.json
timeout 900 python bench.py --workload level2 --steps 3 --no-extras 2>gpurun_out/${T}_bench_level2.err | tail -1 > gpurun_out/${T}_bench_level2.json
for WL in level1 level2; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${T}_${WL}_launches.csv \
      python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --workload $WL > gpurun_out/${T}_${WL}_launches_bench.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:blind_rotate_k1q -s 4 -c 4 -f -o gpurun_out/${T}_${WL}_k1q \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras --workload $WL > gpurun_out/${T}_${WL}_k1q.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:keyswitch -s 1 -c 1 -f -o gpurun_out/${T}_${WL}_ks \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras --workload $WL > gpurun_out/${T}_${WL}_ks.log 2>&1
  # summarise on the box and drop the reports: gpurun brings back at most 64 MiB
  python scripts/ncu_summary.py gpurun_out/${T}_${WL}_k1q.ncu-rep gpurun_out/${T}_${WL}_k1q_ncu_summary.txt
  python scripts/ncu_summary.py gpurun_out/${T}_${WL}_ks.ncu-rep gpurun_out/${T}_${WL}_ks_ncu_summary.txt
  rm -f gpurun_out/${T}_${WL}_k1q.ncu-rep gpurun_out/${T}_${WL}_ks.ncu-rep
done
python - <<'PY'
import json
for f in ("level1", "level2", "reference"):
    try:
        d = json.load(open(f"gpurun_out/r2z_bench_{f}.json"))
        print(f, d.get("value"), (d.get("roofline") or {}).get("frac"), (d.get("e2e") or {}).get("value"), (d.get("e2e_handles") or {}).get("value"))
    except Exception as e:
        print(f, "unreadable", e)
PY
ls -la gpurun_out | grep r2z
