#!/bin/bash
# session-5 cycle: GPU tests, default bench (e2e with device-side marshalling), warm-cache launch lists (ncu --cache-control none)
TAG=${1:-r01j}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py --no-ref-cuda > gpurun_out/bench_c2_${TAG}.json 2> gpurun_out/bench_c2_${TAG}.err
python -c "
import json; d=json.load(open('gpurun_out/bench_c2_${TAG}.json')); print('C2', d['value'], 'e2e', d['e2e']['value'], d['kernels_ms'], d['roofline']['frac'], d['roofline_fp32']['frac'])"
python bench.py --workload c4 --md-steps 200 --steps 3 --warmup 3 --equil 1000 --no-ref-cuda --no-cpu-baseline > gpurun_out/bench_c4_${TAG}.json 2> gpurun_out/bench_c4_${TAG}.err
python -c "
import json; d=json.load(open('gpurun_out/bench_c4_${TAG}.json')); print('C4', d['value'], 'e2e', d['e2e']['value'], d['kernels_ms'])"
# warm-cache per-kernel device times: caches are NOT flushed between kernels (C2's state lives in the 126 MB L2 in production)
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 300 -c 700 --csv --log-file gpurun_out/launches_warm_${TAG}c2.csv \
    python bench.py --steps 1 --warmup 1 --md-steps 150 --equil 300 --no-cpu-baseline --no-ref-cuda > gpurun_out/ncu_llw_${TAG}c2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 200 -c 400 --csv --log-file gpurun_out/launches_warm_${TAG}c4.csv \
    python bench.py --workload c4 --steps 1 --warmup 1 --md-steps 40 --equil 60 --no-cpu-baseline --no-ref-cuda > gpurun_out/ncu_llw_${TAG}c4.log 2>&1
ls -la gpurun_out | tail -8
