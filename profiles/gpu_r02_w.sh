#!/bin/bash
# round 2, call W: recovery from work-list segment overflow; full GPU suite; C4 / C2 quick bench
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 ) > gpurun_out/r2w_tests.log 2>&1
tail -4 gpurun_out/r2w_tests.log
Q="--no-cpu-baseline --no-ref-cuda --no-extras"
run() { tag=$1; wl=$2; shift 2
  env "$@" timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 $Q > gpurun_out/r2w_$tag.json 2> gpurun_out/r2w_$tag.err
  python - <<PY
import json
try:
    b=json.load(open("gpurun_out/r2w_$tag.json")); k=b["kernels_ms"]; print("r2w_$tag", "%.4g" % b["value"], {x: round(k[x],4) for x in ("force_pass","integrate","list_build_per_rebuild","sort_per_sort","md_step_mean")})
except Exception as e: print("r2w_$tag", "failed", e)
PY
}
run c4 c4 X=0
run c2 c2 X=0
run c4_tiny c4 OXB_SEG_SCALE=0.05
