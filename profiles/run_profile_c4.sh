#!/bin/bash
# C4-scale (1M nt) profiling pass: launch list + full capture of the force-pass kernels
set -x
mkdir -p gpurun_out
TAG=${1:-r01c4}
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 300 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --workload c4 --steps 1 --warmup 1 --md-steps 40 --equil 60 --no-cpu-baseline --no-ref-cuda > gpurun_out/ncu_bench_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_edge|k_bonded|k_dh|k_integrate" -s 80 -c 7 -o gpurun_out/prof_forces_${TAG} -f \
    python bench.py --workload c4 --steps 1 --warmup 1 --md-steps 20 --equil 40 --no-cpu-baseline --no-ref-cuda > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out
