#!/bin/bash
# Profiling pass of round 1 (run under gpurun on one B200).  Outputs land in gpurun_out/; summaries are copied to profiles/.
set -x
mkdir -p gpurun_out
TAG=${1:-r01}
# 1. every launch with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 600 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 1 --md-steps 150 --equil 300 --no-cpu-baseline --no-ref-cuda > gpurun_out/ncu_bench_${TAG}.log 2>&1
# 2. full capture of the force kernels
ncu --set full --clock-control none --import-source on -k regex:"k_edge|k_bonded" -s 100 -c 10 -o gpurun_out/prof_forces_${TAG} -f \
    python bench.py --steps 1 --warmup 1 --md-steps 60 --equil 200 --no-cpu-baseline --no-ref-cuda > gpurun_out/ncu_full_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_integrate -s 40 -c 2 -o gpurun_out/prof_integrate_${TAG} -f \
    python bench.py --steps 1 --warmup 1 --md-steps 60 --equil 200 --no-cpu-baseline --no-ref-cuda >> gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out
