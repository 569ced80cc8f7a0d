#!/bin/bash
# round-1 final acceptance cycle (session 5): GPU tests, smoke, default bench (C2) with all legs, reference arm, C4, C3, float precision
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time python bench.py > gpurun_out/bench_c2_final.json 2> gpurun_out/bench_c2_final.err ) 2>&1 | grep real
python -c "
import json; d=json.load(open('gpurun_out/bench_c2_final.json')); print('C2', d['value'], 'e2e', d['e2e']['value'], 'refcuda', (d.get('reference_cuda') or {}).get('value'), 'cpu', (d.get('cpu_baseline') or {}).get('value'), d['kernels_ms'], d['roofline']['frac'], d['roofline_fp32']['frac'], d['roofline']['traffic'], d['clocks'])"
( time python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_arm_final.json 2> gpurun_out/bench_ref_arm_final.err ) 2>&1 | grep real
cut -c1-200 gpurun_out/bench_ref_arm_final.json
python bench.py --workload c4 --md-steps 200 --steps 3 --warmup 3 --equil 1000 --ref-cuda-steps 1000 2000 > gpurun_out/bench_c4_final.json 2> gpurun_out/bench_c4_final.err
python -c "
import json; d=json.load(open('gpurun_out/bench_c4_final.json')); print('C4', d['value'], 'e2e', d['e2e']['value'], 'refcuda', (d.get('reference_cuda') or {}).get('value'), d['kernels_ms'])"
python bench.py --workload c3 --no-cpu-baseline > gpurun_out/bench_c3_final.json 2> gpurun_out/bench_c3_final.err
python -c "
import json; d=json.load(open('gpurun_out/bench_c3_final.json')); print('C3', d['value'], 'e2e', d['e2e']['value'], 'refcuda', (d.get('reference_cuda') or {}).get('value'), d['kernels_ms'])"
