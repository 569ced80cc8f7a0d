#!/bin/bash
# round 2, call K: whole-sector flush of the staged rows (sort / gather pass / neighbour scan / edge fill)
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -x -k "verlet or pair_set or overflow or full_size or forces_torques or plugin or views" 2>&1 | tail -3 ) > gpurun_out/r2k_tests.log 2>&1
tail -1 gpurun_out/r2k_tests.log
Q="--no-cpu-baseline --no-ref-cuda --no-extras"
run() { tag=$1; wl=$2; shift 2
  env "$@" timeout 600 python bench.py --workload $wl --steps 3 --warmup 3 $Q > gpurun_out/r2k_$tag.json 2> gpurun_out/r2k_$tag.err
  python - <<PY
import json
try:
    b=json.load(open("gpurun_out/r2k_$tag.json")); k=b["kernels_ms"]; print("r2k_$tag", "%.4g" % b["value"], {x: round(k[x],4) for x in ("force_pass","integrate","list_build_per_rebuild","sort_per_sort","md_step_mean")}, {x: round(v,4) for x,v in k["rebuild_parts"].items()})
except Exception as e: print("r2k_$tag", "failed", e)
PY
}
run c4 c4 X=0
run c4_full c4 OXB_HALF_SHELL=0
run c2 c2 X=0
