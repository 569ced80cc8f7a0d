#!/bin/bash
# round 2, call AH: oxDNA3 bonds in their own launch; register cap sweep of the neighbour-pass kernel (OXB_DNA3_MB = 3, 4, 5)
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_dna3.py -q -x 2>&1 | tail -5 ) > gpurun_out/r2ah_tests.log 2>&1
tail -1 gpurun_out/r2ah_tests.log
Q="--no-cpu-baseline --no-extras --no-ref-cuda"
for mb in 3 4 5; do
  OXB_DNA3_MB=$mb timeout 600 python bench.py --workload c2_dna3 --steps 3 --warmup 3 $Q > gpurun_out/r2ah_mb$mb.json 2> gpurun_out/r2ah_mb$mb.err
  python - <<PY
import json
try:
    b=json.loads(open("gpurun_out/r2ah_mb$mb.json").read().strip().splitlines()[-1]); k=b.get("kernels_ms")
    print("r2ah_mb$mb", "%.4g" % b["value"], {x: round(k[x],4) for x in ("force_pass","integrate","list_build_per_rebuild","sort_per_sort","md_step_mean")})
except Exception as e: print("r2ah_mb$mb failed", e)
PY
done
