#!/bin/bash
# round 2, call AP: initcheck details (which kernel reads uninitialised device memory?) for an oxDNA3 and an oxDNA2 force test
mkdir -p gpurun_out
( timeout 600 compute-sanitizer --tool initcheck --print-limit 40 python -m pytest tests/test_gpu_dna3.py -q -x -k "forces_torques and lattice8 and mixed and 0-0" 2>&1 | grep -E "Uninitialized|at .*\+0x|by thread|ERROR SUMMARY|passed|failed" | awk '!seen[$0]++' | head -60 ) > gpurun_out/r2ap_init_dna3.log 2>&1
( timeout 600 compute-sanitizer --tool initcheck --print-limit 40 python -m pytest tests/test_gpu_parity.py -q -x -k "forces_torques_energy_vs_reference and lattice8 and 0-0" 2>&1 | grep -E "Uninitialized|at .*\+0x|by thread|ERROR SUMMARY|passed|failed" | awk '!seen[$0]++' | head -60 ) > gpurun_out/r2ap_init_dna2.log 2>&1
tail -3 gpurun_out/r2ap_init_dna3.log; tail -3 gpurun_out/r2ap_init_dna2.log
