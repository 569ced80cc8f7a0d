#!/bin/bash
# round 2, call AC: integrator with every load issued before the first store
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -x -k "nve or thermostat or equipartition or operator or barostat or fix_diffusion or replica or long_run" 2>&1 | tail -3 ) > gpurun_out/r2ac_tests.log 2>&1
tail -1 gpurun_out/r2ac_tests.log
Q="--no-cpu-baseline --no-ref-cuda --no-extras"
run() { tag=$1; wl=$2; shift 2
  env "$@" timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 $Q > gpurun_out/r2ac_$tag.json 2> gpurun_out/r2ac_$tag.err
  python - <<PY
import json
try:
    b=json.load(open("gpurun_out/r2ac_$tag.json")); k=b["kernels_ms"]; print("r2ac_$tag", "%.4g" % b["value"], {x: round(k[x],4) for x in ("force_pass","integrate","list_build_per_rebuild","sort_per_sort","md_step_mean")})
except Exception as e: print("r2ac_$tag", "failed", e)
PY
}
run c4 c4 X=0
run c2 c2 X=0
run c3 c3 X=0
