#!/bin/bash
# round 2, call AJ: cost-split particle-centric kernel for oxDNA2 / oxRNA2 (use_edge = 0): full GPU test suite + A/B against the single loop
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 ) > gpurun_out/r2aj_tests.log 2>&1
tail -1 gpurun_out/r2aj_tests.log
Q="--no-cpu-baseline --no-extras --no-ref-cuda --use-edge 0"
run() { tag=$1; wl=$2; shift 2
  env "$@" timeout 600 python bench.py --workload $wl --steps 3 --warmup 3 $Q > gpurun_out/r2aj_$tag.json 2> gpurun_out/r2aj_$tag.err
  python - <<PY
import json
try:
    b=json.loads(open("gpurun_out/r2aj_$tag.json").read().strip().splitlines()[-1]); k=b.get("kernels_ms")
    print("r2aj_$tag", "%.4g" % b["value"], {x: round(k[x],4) for x in ("force_pass","integrate","list_build_per_rebuild","sort_per_sort","md_step_mean")})
except Exception as e: print("r2aj_$tag failed", e)
PY
}
run c2_single c2 OXB_PARTICLE_SPLIT=0
run c2_split3 c2 OXB_PARTICLE_SPLIT=1 OXB_PARTICLE_MB=3
run c2_split4 c2 OXB_PARTICLE_SPLIT=1 OXB_PARTICLE_MB=4
run c3_single c3 OXB_PARTICLE_SPLIT=0
run c3_split3 c3 OXB_PARTICLE_SPLIT=1 OXB_PARTICLE_MB=3
run c4_single c4 OXB_PARTICLE_SPLIT=0
run c4_split3 c4 OXB_PARTICLE_SPLIT=1 OXB_PARTICLE_MB=3
