#!/bin/bash
mkdir -p gpurun_out
{
for f in 1.0 0.9 1.1 1.0; do
  OXB_SPEC_FACTOR=$f python bench.py --md-steps 1000 --steps 4 --warmup 3 --no-ref-cuda --no-cpu-baseline > gpurun_out/sp.json 2> gpurun_out/sp.err
  python -c "
import json; d=json.load(open('gpurun_out/sp.json')); print('spec_factor = $f  c2 value %.4e step %.4f' % (d['value'], d['kernels_ms']['md_step_mean']))"
done
for f in 0.8 1.0 1.1; do
  OXB_SPEC_FACTOR=$f python bench.py --workload c4 --md-steps 200 --steps 3 --warmup 3 --equil 1000 --no-ref-cuda --no-cpu-baseline > gpurun_out/sp.json 2> gpurun_out/sp.err
  python -c "
import json; d=json.load(open('gpurun_out/sp.json')); print('spec_factor = $f  c4 value %.4e step %.4f' % (d['value'], d['kernels_ms']['md_step_mean']))"
done
} 2>&1 | tee -a gpurun_out/spec_sweep.log
