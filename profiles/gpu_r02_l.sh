#!/bin/bash
# round 2, call L: ncu --set full of the half-shell neighbour kernel at C4 (thermalised state)
mkdir -p gpurun_out
Q="--no-cpu-baseline --no-ref-cuda --no-extras"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_build_neigh|k_fill_edges" -s 20 -c 2 -o gpurun_out/prof_build_half_r02l -f \
    python bench.py --workload c4 --steps 1 --warmup 1 --md-steps 30 --equil 400 $Q > gpurun_out/ncu_build_half_r02l.log 2>&1
ls -la gpurun_out/*r02l*
