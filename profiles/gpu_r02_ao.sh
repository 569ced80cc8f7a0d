#!/bin/bash
# round 2, call AO: compute-sanitizer over the kernels added late in the round: oxDNA3 (k_forces_dna3, k_energy_split_dna3) and the cost-split
# particle-centric kernel (use_edge = 0 force tests, replica batches), memcheck + racecheck + initcheck on the small fixtures
mkdir -p gpurun_out
K="(dna3 and not full_size and not stock_input) or (forces_torques_energy_vs_reference and lattice8) or rna_forces or replica_batch or batch_forces"
( timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --target-processes all python -m pytest tests -m gpu -q -x -k "$K" 2>&1 | tail -12 ) > gpurun_out/r2ao_memcheck.log 2>&1
tail -4 gpurun_out/r2ao_memcheck.log
( timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_dna3.py -q -x -k "forces_torques and lattice8 and mixed" 2>&1 | tail -8 ) > gpurun_out/r2ao_racecheck.log 2>&1
tail -3 gpurun_out/r2ao_racecheck.log
( timeout 900 compute-sanitizer --tool initcheck --error-exitcode 9 python -m pytest tests/test_gpu_dna3.py -q -x -k "forces_torques and lattice8 and mixed or nicked" 2>&1 | tail -8 ) > gpurun_out/r2ao_initcheck.log 2>&1
tail -3 gpurun_out/r2ao_initcheck.log
( timeout 600 python -m pytest tests/test_gpu_dna3.py -q -x -k "refuses" 2>&1 | tail -3 ) > gpurun_out/r2ao_new.log 2>&1
tail -1 gpurun_out/r2ao_new.log
