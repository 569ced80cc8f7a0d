#!/bin/bash
# round 2, call AK: half shell by cell direction (13 forward cells) instead of by slot order: balanced rows for the list build, the Debye-Hueckel
# kernel and the near-edge segments (OXB_GEO_HALF = 0 / 1)
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q -x -k "not dropin" 2>&1 | tail -5 ) > gpurun_out/r2ak_tests.log 2>&1
tail -1 gpurun_out/r2ak_tests.log
Q="--no-cpu-baseline --no-ref-cuda --no-extras"
run() { tag=$1; wl=$2; shift 2
  env "$@" timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 $Q $EXTRA > gpurun_out/r2ak_$tag.json 2> gpurun_out/r2ak_$tag.err
  python - <<PY
import json
try:
    b=json.loads(open("gpurun_out/r2ak_$tag.json").read().strip().splitlines()[-1]); k=b.get("kernels_ms")
    print("r2ak_$tag", "%.4g" % b["value"], {x: round(k[x],4) for x in ("force_pass","integrate","list_build_per_rebuild","sort_per_sort","md_step_mean")} if k else "")
except Exception as e: print("r2ak_$tag", "failed", e)
PY
}
run c4_geo c4 OXB_GEO_HALF=1
run c4_slot c4 OXB_GEO_HALF=0
run c2_geo c2 OXB_GEO_HALF=1
run c2_slot c2 OXB_GEO_HALF=0
run c3_geo c3 OXB_GEO_HALF=1
run c3_slot c3 OXB_GEO_HALF=0
EXTRA="--replicas 8" run c5_geo c5 OXB_GEO_HALF=1
EXTRA="--replicas 8" run c5_slot c5 OXB_GEO_HALF=0
