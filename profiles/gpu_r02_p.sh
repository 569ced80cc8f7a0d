#!/bin/bash
# round 2, call P: one-launch ordering of small systems (cooperative counting sort); full GPU suite
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15 ) > gpurun_out/r2p_tests.log 2>&1
tail -4 gpurun_out/r2p_tests.log
Q="--no-cpu-baseline --no-ref-cuda --no-extras"
run() { tag=$1; wl=$2; shift 2
  env "$@" timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 $Q > gpurun_out/r2p_$tag.json 2> gpurun_out/r2p_$tag.err
  python - <<PY
import json
try:
    b=json.load(open("gpurun_out/r2p_$tag.json")); k=b["kernels_ms"]; print("r2p_$tag", "%.4g" % b["value"], {x: round(k[x],4) for x in ("force_pass","integrate","list_build_per_rebuild","sort_per_sort","md_step_mean")}, {x: round(v,4) for x,v in k["rebuild_parts"].items()})
except Exception as e: print("r2p_$tag", "failed", e)
PY
}
run c2 c2 X=0
run c2_cub c2 OXB_SORT_SMALL=0
run c3 c3 X=0
run small small X=0
