#!/bin/bash
# length of the speculative launch batch (fraction of the running mean rebuild interval) at C2
mkdir -p gpurun_out
for f in 0.8 0.6 1.0 1.2 0.8; do
  OXB_SPEC_FACTOR=$f python bench.py --md-steps 1000 --steps 4 --warmup 3 --no-ref-cuda --no-cpu-baseline > gpurun_out/sp.json 2> gpurun_out/sp.err
  python -c "
import json; d=json.load(open('gpurun_out/sp.json')); print('spec_factor = $f  c2 value %.4e step %.4f' % (d['value'], d['kernels_ms']['md_step_mean']))"
done 2>&1 | tee gpurun_out/spec_sweep.log
