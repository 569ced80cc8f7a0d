#!/bin/bash
# round 2, call AA: software-pipelined L1 prefetch of the pair kernels' gathers (OXB_PREFETCH=0/1)
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -x -k "forces_torques or rna_forces or full_size or nve or replica or work_list" 2>&1 | tail -3 ) > gpurun_out/r2aa_tests.log 2>&1
tail -1 gpurun_out/r2aa_tests.log
Q="--no-cpu-baseline --no-ref-cuda --no-extras"
run() { tag=$1; wl=$2; shift 2
  env "$@" timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 $Q > gpurun_out/r2aa_$tag.json 2> gpurun_out/r2aa_$tag.err
  python - <<PY
import json
try:
    b=json.load(open("gpurun_out/r2aa_$tag.json")); k=b["kernels_ms"]; print("r2aa_$tag", "%.4g" % b["value"], {x: round(k[x],4) for x in ("force_pass","integrate","list_build_per_rebuild","sort_per_sort","md_step_mean")})
except Exception as e: print("r2aa_$tag", "failed", e)
PY
}
run c4_pf c4 OXB_PREFETCH=1
run c4_nopf c4 OXB_PREFETCH=0
run c2_pf c2 OXB_PREFETCH=1
run c2_nopf c2 OXB_PREFETCH=0
run c3_pf c3 OXB_PREFETCH=1
run c3_nopf c3 OXB_PREFETCH=0
