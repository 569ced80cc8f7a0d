#!/bin/bash
# quick launch lists (device time per kernel) for C2 and C4; usage: launchlist.sh <tag>
TAG=${1:-cur}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 700 --csv --log-file gpurun_out/launches_${TAG}c2.csv \
    python bench.py --steps 1 --warmup 1 --md-steps 150 --equil 300 --no-cpu-baseline --no-ref-cuda > gpurun_out/ncu_ll_${TAG}c2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/launches_${TAG}c4.csv \
    python bench.py --workload c4 --steps 1 --warmup 1 --md-steps 40 --equil 60 --no-cpu-baseline --no-ref-cuda > gpurun_out/ncu_ll_${TAG}c4.log 2>&1
