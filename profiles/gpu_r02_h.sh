#!/bin/bash
# round 2, call H: balanced DH rows under the half shell; co-resident launch grids of the force pass (OXB_PB_* = blocks per SM)
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15 ) > gpurun_out/r2h_tests.log 2>&1
tail -4 gpurun_out/r2h_tests.log
Q="--no-cpu-baseline --no-ref-cuda --no-extras"
run() { # tag workload env...
  tag=$1; wl=$2; shift 2
  env "$@" timeout 600 python bench.py --workload $wl --steps 3 --warmup 3 $Q > gpurun_out/r2h_$tag.json 2> gpurun_out/r2h_$tag.err
  python - <<PY
import json
try:
    b=json.load(open("gpurun_out/r2h_$tag.json")); k=b["kernels_ms"]; print("r2h_$tag", "%.4g" % b["value"], {x: round(k[x],4) for x in ("force_pass","integrate","list_build_per_rebuild","sort_per_sort","md_step_mean")})
except Exception as e: print("r2h_$tag", "failed", e)
PY
}
run c4_base c4 X=0
run c4_p1 c4 OXB_PB_NEAR=4 OXB_PB_DH=3 OXB_PB_BONDED=2
run c4_p2 c4 OXB_PB_NEAR=6 OXB_PB_DH=4 OXB_PB_BONDED=2
run c4_p3 c4 OXB_PB_NEAR=8 OXB_PB_DH=4 OXB_PB_BONDED=3
run c4_p4 c4 OXB_PB_NEAR=16 OXB_PB_DH=4 OXB_PB_BONDED=2
run c4_p5 c4 OXB_PB_NEAR=8 OXB_PB_DH=6 OXB_PB_BONDED=2
run c2_base c2 X=0
run c2_p1 c2 OXB_PB_NEAR=4 OXB_PB_DH=3 OXB_PB_BONDED=2
run c2_p3 c2 OXB_PB_NEAR=8 OXB_PB_DH=4 OXB_PB_BONDED=3
