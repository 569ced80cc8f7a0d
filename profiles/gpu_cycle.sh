python profiles/error_report.py 2>&1 | tail -6
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --no-cpu-baseline --no-ref-cuda > gpurun_out/bench_c2_cur.json 2> gpurun_out/bench_cur.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c2_cur.json')); print('C2', d['value'], d['kernels_ms'], d['e2e']['value'], d['gpu_launches'], d['config']['list_rebuild_every_md_steps'])"; tail -3 gpurun_out/bench_cur.err
python bench.py --workload c4 --md-steps 200 --steps 3 --warmup 3 --equil 1000 --no-cpu-baseline --no-ref-cuda > gpurun_out/bench_c4_cur.json 2> gpurun_out/bench_c4.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c4_cur.json')); print('C4', d['value'], d['kernels_ms'], d['e2e']['value'], d['config']['list_rebuild_every_md_steps'])"; tail -3 gpurun_out/bench_c4.err
