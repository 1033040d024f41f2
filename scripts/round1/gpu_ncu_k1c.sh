#!/usr/bin/env bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:blind_rotate_k1c -s 2 -c 1 -f -o gpurun_out/k1c_prof \
    python scripts/latency_breakdown.py > gpurun_out/k1c_prof.log 2>&1
python scripts/ncu_summary.py gpurun_out/k1c_prof.ncu-rep > gpurun_out/k1c_ncu_summary.txt 2>&1
rm -f gpurun_out/k1c_prof.ncu-rep
cat gpurun_out/k1c_ncu_summary.txt
