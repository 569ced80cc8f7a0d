#!/bin/bash
# round 2, call AE: oxDNA3 on the GPU (tests/test_gpu_dna3.py) + the regression subset of the other GPU tests
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_dna3.py -q -x 2>&1 | tail -40 ) > gpurun_out/r2ae_dna3.log 2>&1
tail -5 gpurun_out/r2ae_dna3.log
( timeout 900 python -m pytest tests -m gpu -q -x -k "forces_torques or rna_forces or nve or tiny or plugin" 2>&1 | tail -3 ) > gpurun_out/r2ae_tests.log 2>&1
tail -1 gpurun_out/r2ae_tests.log
