#!/bin/bash
# round 2, call AD: row-major Debye-Hueckel matrix + flat deal of a warp's pairs to its lanes (OXB_DH_FLAT=0/1)
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -x -k "forces_torques or rna_forces or full_size or nve or replica or work_list or first_generation or tiny or single" 2>&1 | tail -3 ) > gpurun_out/r2ad_tests.log 2>&1
tail -1 gpurun_out/r2ad_tests.log
Q="--no-cpu-baseline --no-ref-cuda --no-extras"
run() { tag=$1; wl=$2; shift 2
  env "$@" timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 $Q $EXTRA > gpurun_out/r2ad_$tag.json 2> gpurun_out/r2ad_$tag.err
  python - <<PY
import json
try:
    b=json.loads(open("gpurun_out/r2ad_$tag.json").read().strip().splitlines()[-1]); k=b.get("kernels_ms")
    print("r2ad_$tag", "%.4g" % b["value"], {x: round(k[x],4) for x in ("force_pass","integrate","list_build_per_rebuild","sort_per_sort","md_step_mean")} if k else "")
except Exception as e: print("r2ad_$tag", "failed", e)
PY
}
run c4_flat c4 OXB_DH_FLAT=1
run c4_rows c4 OXB_DH_FLAT=0
run c2_flat c2 OXB_DH_FLAT=1
run c2_rows c2 OXB_DH_FLAT=0
run c3_flat c3 OXB_DH_FLAT=1
EXTRA="--replicas 8" run c5_flat c5 OXB_DH_FLAT=1
EXTRA="--replicas 8" run c5_rows c5 OXB_DH_FLAT=0
