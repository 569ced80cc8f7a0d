#!/bin/bash
# round 2, call AV: ncu of the oxDNA3 edge-pipeline kernels at 1M nucleotides
mkdir -p gpurun_out
Q="--no-cpu-baseline --no-extras --no-ref-cuda"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k3_|k_dh_particle" -s 300 -c 5 -o gpurun_out/prof_dna3e_c4_r02av -f python bench.py --workload c4_dna3 --steps 1 --warmup 1 --md-steps 30 --equil 300 $Q > gpurun_out/ncu_dna3e_c4_r02av.log 2>&1
tail -1 gpurun_out/ncu_dna3e_c4_r02av.log | cut -c1-60
