#!/bin/bash
# round 2, call AX: oxDNA3 with dummy bases and custom base types on the GPU (both force variants)
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_dna3.py -q -x 2>&1 | tail -12 ) > gpurun_out/r2ax_tests.log 2>&1
tail -3 gpurun_out/r2ax_tests.log
