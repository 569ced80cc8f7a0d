#!/bin/bash
# round 2, call D: group-per-particle neighbour scan (OXB_BUILD_G sweep), cell table + staleness refs folded into the re-sort's gather pass
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 ) > gpurun_out/r2d_tests.log 2>&1
tail -3 gpurun_out/r2d_tests.log
Q="--no-cpu-baseline --no-ref-cuda --no-extras"
for G in 8 1 4 16; do
  OXB_BUILD_G=$G timeout 600 python bench.py --workload c4 --steps 3 --warmup 3 $Q > gpurun_out/r2d_c4_g$G.json 2> gpurun_out/r2d_c4_g$G.err
done
OXB_BUILD_G=8 timeout 600 python bench.py --workload c2 --steps 5 --warmup 3 $Q > gpurun_out/r2d_c2_g8.json 2> gpurun_out/r2d_c2_g8.err
for f in r2d_c4_g8 r2d_c4_g1 r2d_c4_g4 r2d_c4_g16 r2d_c2_g8; do python - <<PY
import json
try:
    b=json.load(open("gpurun_out/$f.json")); k=b["kernels_ms"]; print("$f", "%.4g" % b["value"], {x: round(k[x],4) for x in ("force_pass","integrate","list_build_per_rebuild","sort_per_sort","md_step_mean")})
except Exception as e: print("$f", "failed", e)
PY
done
