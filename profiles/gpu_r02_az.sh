#!/bin/bash
# round 2, call AZ: ncu launch lists (gpu__time_duration.sum, cold cache, serialised: compare SHARES) of the bench command on the final code: C4 and oxDNA3 on the C2 geometry
mkdir -p gpurun_out
Q="--no-cpu-baseline --no-ref-cuda --no-extras"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 40000 -c 900 --csv --log-file gpurun_out/launches_c4_r02e.csv \
  python bench.py --workload c4 --steps 1 --warmup 1 --equil 6000 --md-steps 100 $Q > gpurun_out/ncu_launches_c4_r02e.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 30000 -c 900 --csv --log-file gpurun_out/launches_c2_dna3_r02e.csv \
  python bench.py --workload c2_dna3 --steps 1 --warmup 1 --equil 6000 --md-steps 100 $Q > gpurun_out/ncu_launches_c2_dna3_r02e.log 2>&1
ls -la gpurun_out/launches_*r02e.csv
