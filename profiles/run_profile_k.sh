#!/bin/bash
# Profiling pass r01k (one B200 under gpurun): launch list of the default bench command (cold-cache, serialised => compare SHARES) and
# ncu --set full captures of every kernel of the force pass (Debye-Hueckel included) and of the fused integrator, C2 and C4.
set -x
mkdir -p gpurun_out
TAG=${1:-r01k}
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 700 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 1 --md-steps 150 --equil 300 --no-cpu-baseline --no-ref-cuda > gpurun_out/ncu_bench_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_edge|k_bonded|k_dh|k_excl" -s 120 -c 12 -o gpurun_out/prof_forces_${TAG} -f \
    python bench.py --steps 1 --warmup 1 --md-steps 60 --equil 200 --no-cpu-baseline --no-ref-cuda > gpurun_out/ncu_full_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_integrate -s 40 -c 2 -o gpurun_out/prof_integrate_${TAG} -f \
    python bench.py --steps 1 --warmup 1 --md-steps 60 --equil 200 --no-cpu-baseline --no-ref-cuda >> gpurun_out/ncu_full_${TAG}.log 2>&1
# C4 (1M nucleotides)
ncu --set full --clock-control none --import-source on -k regex:"k_edge|k_bonded|k_dh|k_excl|k_integrate" -s 70 -c 7 -o gpurun_out/prof_c4_${TAG} -f \
    python bench.py --workload c4 --steps 1 --warmup 1 --md-steps 30 --equil 60 --no-cpu-baseline --no-ref-cuda >> gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out | tail -12
