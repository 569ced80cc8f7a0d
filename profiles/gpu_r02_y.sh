#!/bin/bash
# round 2, call Y: compute-sanitizer memcheck over the kernels that are new this round (small fixtures), both neighbour kernels
mkdir -p gpurun_out
K="verlet_pair_set or forces_torques_energy_vs_reference or work_list_segment or toy_plugin or device_views or batch_forces or one_launch or rna_forces"
( timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --target-processes all python -m pytest tests -m gpu -q -x -k "$K" 2>&1 | tail -25 ) > gpurun_out/r2y_memcheck.log 2>&1
tail -6 gpurun_out/r2y_memcheck.log
( OXB_BUILD_G=1 timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q -x -k "verlet_pair_set or forces_torques_energy_vs_reference" 2>&1 | tail -8 ) > gpurun_out/r2y_memcheck_g1.log 2>&1
tail -4 gpurun_out/r2y_memcheck_g1.log
( timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -q -x -k "forces_torques_energy_vs_reference and lattice8" 2>&1 | tail -8 ) > gpurun_out/r2y_racecheck.log 2>&1
tail -4 gpurun_out/r2y_racecheck.log
