#!/bin/bash
# round 2, call T: blocks per SM of the near-edge kernel at small sizes
mkdir -p gpurun_out
Q="--no-cpu-baseline --no-ref-cuda --no-extras"
run() { tag=$1; wl=$2; shift 2
  env "$@" timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 $Q > gpurun_out/r2t_$tag.json 2> gpurun_out/r2t_$tag.err
  python - <<PY
import json
try:
    b=json.load(open("gpurun_out/r2t_$tag.json")); k=b["kernels_ms"]; print("r2t_$tag", "%.4g" % b["value"], {x: round(k[x],4) for x in ("force_pass","integrate","list_build_per_rebuild","sort_per_sort","md_step_mean")})
except Exception as e: print("r2t_$tag", "failed", e)
PY
}
run c2_16 c2 X=0
run c2_24 c2 OXB_PB_NEAR=24
run c2_32 c2 OXB_PB_NEAR=32
run c2_8 c2 OXB_PB_NEAR=8
run c3_24 c3 OXB_PB_NEAR=24
run c3_32 c3 OXB_PB_NEAR=32
