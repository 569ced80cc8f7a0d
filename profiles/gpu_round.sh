# round-1 acceptance cycle: GPU tests, smoke, default bench (C2) with all legs, reference arm, C4 and C3 benches
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time python bench.py > gpurun_out/bench_c2_full.json 2> gpurun_out/bench_c2_full.err ) 2>&1 | grep real
python -c "
import json; d=json.load(open('gpurun_out/bench_c2_full.json')); print('C2', d['value'], 'e2e', d['e2e']['value'], 'refcuda', (d.get('reference_cuda') or {}).get('value'), 'cpu', (d.get('cpu_baseline') or {}).get('value'), d['kernels_ms'], d['roofline']['frac'], d['roofline_fp32']['frac'], d['clocks'])"
( time python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_arm.json 2> gpurun_out/bench_ref_arm.err ) 2>&1 | grep real
cut -c1-300 gpurun_out/bench_ref_arm.json
python bench.py --workload c4 --md-steps 200 --steps 3 --warmup 3 --equil 1000 --ref-cuda-steps 1000 2000 > gpurun_out/bench_c4_full.json 2> gpurun_out/bench_c4_full.err
python -c "
import json; d=json.load(open('gpurun_out/bench_c4_full.json')); print('C4', d['value'], 'e2e', d['e2e']['value'], 'refcuda', (d.get('reference_cuda') or {}).get('value'), d['kernels_ms'])"
