#!/bin/bash
# round 2, call AS: drop-in oxDNA3 runs (use_edge = 1 now the staged pipeline), memcheck + racecheck of the k3_* kernels, ncu of the oxDNA3 force pass
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_dropin.py -q -x -k "dna3" 2>&1 | tail -4 ) > gpurun_out/r2as_dropin.log 2>&1
tail -1 gpurun_out/r2as_dropin.log
( timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_dna3.py -q -x -k "(forces_torques and 1-1) or nicked or model_switch" 2>&1 | tail -4 ) > gpurun_out/r2as_memcheck.log 2>&1
tail -2 gpurun_out/r2as_memcheck.log
( timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_dna3.py -q -x -k "forces_torques and lattice8 and mixed and 1-1" 2>&1 | tail -4 ) > gpurun_out/r2as_racecheck.log 2>&1
tail -2 gpurun_out/r2as_racecheck.log
Q="--no-cpu-baseline --no-extras --no-ref-cuda"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k3_|k_dh_particle" -s 400 -c 10 -o gpurun_out/prof_dna3e_r02as -f python bench.py --workload c2_dna3 --steps 1 --warmup 1 --md-steps 60 --equil 400 $Q > gpurun_out/ncu_dna3e_r02as.log 2>&1
tail -1 gpurun_out/ncu_dna3e_r02as.log | cut -c1-60
