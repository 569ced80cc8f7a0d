#!/bin/bash
# round 2, call BA: half-matrix Debye-Hueckel kernel with lane pairs chosen by row length (OXB_DH_SORTED = 1 / 0)
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -x -k "forces_torques or rna_forces or full_size or nve or replica or dna3 or tiny" 2>&1 | tail -3 ) > gpurun_out/r2ba_tests.log 2>&1
tail -1 gpurun_out/r2ba_tests.log
Q="--no-cpu-baseline --no-ref-cuda --no-extras"
run() { tag=$1; wl=$2; shift 2
  env "$@" timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 $Q $EXTRA > gpurun_out/r2ba_$tag.json 2> gpurun_out/r2ba_$tag.err
  python - <<PY
import json
try:
    b=json.loads(open("gpurun_out/r2ba_$tag.json").read().strip().splitlines()[-1]); k=b.get("kernels_ms")
    print("r2ba_$tag", "%.4g" % b["value"], {x: round(k[x],4) for x in ("force_pass","integrate","list_build_per_rebuild","sort_per_sort","md_step_mean")} if k else "")
except Exception as e: print("r2ba_$tag", "failed", e)
PY
}
run c4_sorted c4 OXB_DH_SORTED=1
run c4_fixed c4 OXB_DH_SORTED=0
run c2_sorted c2 OXB_DH_SORTED=1
run c2_fixed c2 OXB_DH_SORTED=0
EXTRA="--replicas 8" run c5_sorted c5 OXB_DH_SORTED=1
EXTRA="--replicas 8" run c5_fixed c5 OXB_DH_SORTED=0
