#!/usr/bin/env bash
# round 2, call I: segmented blind rotation at level 2 (key > L2): time and DRAM traffic with / without
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "segmented or full_size_round_trip or fullsize" 2>&1 | tail -4 | tee gpurun_out/r2i_pytest.log
for seg in 0 1; do
  echo "== segments $([ $seg = 1 ] && echo on || echo off)"
  if [ $seg = 0 ]; then export MB200_NO_SEGMENTS=1; else unset MB200_NO_SEGMENTS; fi
  WLS=level2 POLICIES=5 timeout 600 python scripts/k1q_time.py 2>&1 | grep -v "fp64 peak"
done | tee gpurun_out/r2i_segments.log
unset MB200_NO_SEGMENTS
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:blind_rotate_k1q -s 2 -c 2 --csv \
    --log-file gpurun_out/r2i_level2_segments_dram.csv env WLS=level2 POLICIES=5 python scripts/k1q_time.py > /dev/null 2>&1
MB200_NO_SEGMENTS=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:blind_rotate_k1q -s 1 -c 1 --csv \
    --log-file gpurun_out/r2i_level2_single_dram.csv env WLS=level2 POLICIES=5 python scripts/k1q_time.py > /dev/null 2>&1
tail -5 gpurun_out/r2i_level2_segments_dram.csv gpurun_out/r2i_level2_single_dram.csv
