#!/bin/bash
# round 2, call U: blocks per SM of the near-edge kernel (and of its work-list segments): fewer, longer-lived blocks
mkdir -p gpurun_out
Q="--no-cpu-baseline --no-ref-cuda --no-extras"
run() { tag=$1; wl=$2; shift 2
  env "$@" timeout 600 python bench.py --workload $wl --steps 4 --warmup 3 $Q > gpurun_out/r2u_$tag.json 2> gpurun_out/r2u_$tag.err
  python - <<PY
import json
try:
    b=json.load(open("gpurun_out/r2u_$tag.json")); k=b["kernels_ms"]; print("r2u_$tag", "%.4g" % b["value"], {x: round(k[x],4) for x in ("force_pass","integrate","md_step_mean")})
except Exception as e: print("r2u_$tag", "failed", e)
PY
}
for pb in 3 4 6 10 12; do run c2_$pb c2 OXB_PB_NEAR=$pb; done
for pb in 4 6 8; do run c3_$pb c3 OXB_PB_NEAR=$pb; done
for pb in 6 8 12; do run c4_$pb c4 OXB_PB_NEAR=$pb; done
for pb in 4 8; do run small_$pb small OXB_PB_NEAR=$pb; done
