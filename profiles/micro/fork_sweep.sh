#!/bin/bash
# force pass on three concurrent streams (OXB_FORK=1) vs in line on one stream (OXB_FORK=0): C2, C4, C3
mkdir -p gpurun_out
for rep in 1 2; do
for f in 1 0; do
  OXB_FORK=$f python bench.py --workload c4 --md-steps 200 --steps 3 --warmup 3 --equil 600 --no-ref-cuda --no-cpu-baseline > gpurun_out/fk.json 2> gpurun_out/fk.err
  python -c "
import json; d=json.load(open('gpurun_out/fk.json')); print('fork = $f  c4 value %.3e forces_ms %.4f step %.4f' % (d['value'], d['kernels_ms']['forces'], d['kernels_ms']['md_step_mean']))"
  OXB_FORK=$f python bench.py --md-steps 1000 --steps 4 --warmup 3 --no-ref-cuda --no-cpu-baseline > gpurun_out/fk.json 2> gpurun_out/fk.err
  python -c "
import json; d=json.load(open('gpurun_out/fk.json')); print('fork = $f  c2 value %.3e forces_ms %.4f step %.4f' % (d['value'], d['kernels_ms']['forces'], d['kernels_ms']['md_step_mean']))"
done
done 2>&1 | tee gpurun_out/fork_sweep.log
