#!/usr/bin/env bash
# ncu captures on the GPU box (one GPU). $1 = tag for output names, $2 = workload (level1|level2)
set -x
TAG=${1:-r1}; WL=${2:-level1}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_${WL}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --workload $WL > gpurun_out/${TAG}_${WL}_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:blind_rotate_k1 -s 1 -c 1 -f -o gpurun_out/${TAG}_${WL}_k1 \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --workload $WL > gpurun_out/${TAG}_${WL}_k1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:keyswitch -s 1 -c 1 -f -o gpurun_out/${TAG}_${WL}_ks \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --workload $WL > gpurun_out/${TAG}_${WL}_ks.log 2>&1
ls -la gpurun_out
