#!/bin/bash
# round 2, call A: full GPU suite (new replica-batch tests, 1M-nt oracle test) + the new default bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
nproc >> gpurun_out/r2a_smi.txt; free -g >> gpurun_out/r2a_smi.txt
( time timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > gpurun_out/r2a_tests.log 2>&1
( time timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err ) > gpurun_out/r2a_bench.time 2>&1
tail -5 gpurun_out/r2a_tests.log
tail -c 600 gpurun_out/r2a_bench.json
