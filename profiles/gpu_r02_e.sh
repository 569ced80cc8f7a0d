#!/bin/bash
# round 2, call E: neighbour scan without the per-warp same-address atomic, OXB_BUILD_G sweep
mkdir -p gpurun_out
Q="--no-cpu-baseline --no-ref-cuda --no-extras"
( timeout 600 python -m pytest tests -m gpu -q -x -k "verlet or pair_set or overflow or full_size_c2" 2>&1 | tail -3 ) > gpurun_out/r2e_tests.log 2>&1
tail -1 gpurun_out/r2e_tests.log
for G in 8 1 4 16; do
  OXB_BUILD_G=$G timeout 600 python bench.py --workload c4 --steps 3 --warmup 3 $Q > gpurun_out/r2e_c4_g$G.json 2> gpurun_out/r2e_c4_g$G.err
done
for G in 8 1 4; do
OXB_BUILD_G=$G timeout 600 python bench.py --workload c2 --steps 5 --warmup 3 $Q > gpurun_out/r2e_c2_g$G.json 2> gpurun_out/r2e_c2_g$G.err
done
for f in r2e_c4_g8 r2e_c4_g1 r2e_c4_g4 r2e_c4_g16 r2e_c2_g8 r2e_c2_g1 r2e_c2_g4; do python - <<PY
import json
try:
    b=json.load(open("gpurun_out/$f.json")); k=b["kernels_ms"]; print("$f", "%.4g" % b["value"], {x: round(k[x],4) for x in ("force_pass","integrate","list_build_per_rebuild","sort_per_sort","md_step_mean")})
except Exception as e: print("$f", "failed", e)
PY
done
