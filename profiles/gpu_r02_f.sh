#!/bin/bash
# round 2, call F: ncu --set full of the list-build kernels at C4 (thread-per-particle and 8 lanes per particle), thermalised state
mkdir -p gpurun_out
Q="--no-cpu-baseline --no-ref-cuda --no-extras"
for G in 1 8; do
OXB_BUILD_G=$G timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_build_neigh|k_fill_edges|k_permute" -s 30 -c 3 -o gpurun_out/prof_build_g${G}_r02f -f \
    python bench.py --workload c4 --steps 1 --warmup 1 --md-steps 30 --equil 400 $Q > gpurun_out/ncu_build_g${G}_r02f.log 2>&1
done
ls -la gpurun_out/*.ncu-rep | tail -4
