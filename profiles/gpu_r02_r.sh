#!/bin/bash
# round 2, call R: lane pairs share the rows in the classification pass of the serial neighbour kernel (exercised on the small fixtures with
# OXB_BUILD_G=1, at full size by the C4 oracle test), out-of-run overflow check
mkdir -p gpurun_out
( OXB_BUILD_G=1 timeout 900 python -m pytest tests -m gpu -q -x -k "verlet or pair_set or overflow or forces_torques or full_size or replica or rna_forces or views or plugin or one_launch" 2>&1 | tail -3 ) > gpurun_out/r2r_tests_g1.log 2>&1
tail -1 gpurun_out/r2r_tests_g1.log
( OXB_BUILD_G=1 OXB_HALF_SHELL=0 timeout 900 python -m pytest tests -m gpu -q -x -k "verlet or pair_set or forces_torques or full_size_c2" 2>&1 | tail -3 ) > gpurun_out/r2r_tests_g1_full.log 2>&1
tail -1 gpurun_out/r2r_tests_g1_full.log
Q="--no-cpu-baseline --no-ref-cuda --no-extras"
run() { tag=$1; wl=$2; shift 2
  env "$@" timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 $Q $EXTRA > gpurun_out/r2r_$tag.json 2> gpurun_out/r2r_$tag.err
  python - <<PY
import json
try:
    b=json.load(open("gpurun_out/r2r_$tag.json")); k=b["kernels_ms"]; print("r2r_$tag", "%.4g" % b["value"], {x: round(k[x],4) for x in ("force_pass","integrate","list_build_per_rebuild","sort_per_sort","md_step_mean")}, {x: round(v,4) for x,v in k["rebuild_parts"].items()})
except Exception as e: print("r2r_$tag", "failed", e)
PY
}
run c4 c4 X=0
EXTRA="--replicas 8" run c5_8 c5 X=0
