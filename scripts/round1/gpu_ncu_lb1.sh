#!/usr/bin/env bash
mkdir -p gpurun_out
export MB200_K1_LB=1 MB200_K1_PF=0
ncu --set full --clock-control none --import-source on -k regex:blind_rotate_k1 -s 1 -c 1 -f -o gpurun_out/lb1_k1 \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/lb1_k1.log 2>&1
python scripts/ncu_summary.py gpurun_out/lb1_k1.ncu-rep > gpurun_out/lb1_k1_ncu_summary.txt 2>&1
rm -f gpurun_out/lb1_k1.ncu-rep
cat gpurun_out/lb1_k1_ncu_summary.txt
