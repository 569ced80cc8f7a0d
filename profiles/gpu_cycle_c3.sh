# round-1 C3 (oxRNA2) cycle: GPU tests, C3 bench incl. the reference CUDA backend and CPU legs, ncu launch list of the C3 step
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --workload c3 --ref-cuda-steps 5000 10000 > gpurun_out/bench_c3_cur.json 2> gpurun_out/bench_c3.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c3_cur.json')); print('C3', d['value'], d['kernels_ms'], d['e2e']['value'], d['gpu_launches'], d['config']['list_rebuild_every_md_steps'], d['config']['pairs_per_particle']); print('refcuda', (d.get('reference_cuda') or {}).get('best')); print('cpu', d.get('cpu_baseline')); print('U/N', d['e2e']['U_per_particle'], d['e2e']['K_per_particle'])"; tail -3 gpurun_out/bench_c3.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/launches_c3r01.csv \
    python bench.py --workload c3 --steps 1 --warmup 1 --md-steps 150 --equil 300 --no-cpu-baseline --no-ref-cuda > gpurun_out/ncu_bench_c3.log 2>&1
python profiles/summarize.py c3r01; tail -22 profiles/summary_c3r01.txt; cp profiles/summary_c3r01.txt gpurun_out/
