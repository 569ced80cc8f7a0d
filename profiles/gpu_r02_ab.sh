#!/bin/bash
# round 2, call AB: edge list in three class-group segments (OXB_CLASS_GROUPS=0/1)
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -x -k "forces_torques or rna_forces or full_size or nve or replica or work_list or views or overflow or pair_set" 2>&1 | tail -3 ) > gpurun_out/r2ab_tests.log 2>&1
tail -1 gpurun_out/r2ab_tests.log
Q="--no-cpu-baseline --no-ref-cuda --no-extras"
run() { tag=$1; wl=$2; shift 2
  env "$@" timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 $Q > gpurun_out/r2ab_$tag.json 2> gpurun_out/r2ab_$tag.err
  python - <<PY
import json
try:
    b=json.load(open("gpurun_out/r2ab_$tag.json")); k=b["kernels_ms"]; print("r2ab_$tag", "%.4g" % b["value"], {x: round(k[x],4) for x in ("force_pass","integrate","list_build_per_rebuild","sort_per_sort","md_step_mean")})
except Exception as e: print("r2ab_$tag", "failed", e)
PY
}
run c4_cg c4 OXB_CLASS_GROUPS=1
run c4_nocg c4 OXB_CLASS_GROUPS=0
run c2_cg c2 OXB_CLASS_GROUPS=1
run c2_nocg c2 OXB_CLASS_GROUPS=0
run c3_cg c3 OXB_CLASS_GROUPS=1
run c3_nocg c3 OXB_CLASS_GROUPS=0
