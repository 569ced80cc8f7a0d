#!/bin/bash
# round 2, call Q: hydrogen bonding / cross stacking folded into the tail of the near-edge kernel (small systems); quick MD tests with fixed seed
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q -k "quick_md or forces or full_size_c2 or nve or rna" 2>&1 | tail -6 ) > gpurun_out/r2q_tests.log 2>&1
tail -4 gpurun_out/r2q_tests.log
Q="--no-cpu-baseline --no-ref-cuda --no-extras"
run() { tag=$1; wl=$2; shift 2
  env "$@" timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 $Q > gpurun_out/r2q_$tag.json 2> gpurun_out/r2q_$tag.err
  python - <<PY
import json
try:
    b=json.load(open("gpurun_out/r2q_$tag.json")); k=b["kernels_ms"]; print("r2q_$tag", "%.4g" % b["value"], {x: round(k[x],4) for x in ("force_pass","integrate","list_build_per_rebuild","sort_per_sort","md_step_mean")})
except Exception as e: print("r2q_$tag", "failed", e)
PY
}
run c2 c2 X=0
run c2_nofold c2 OXB_FOLD_HB=0
run c3 c3 X=0
run c3_nofold c3 OXB_FOLD_HB=0
run c4_fold c4 OXB_FOLD_HB=1
run small small X=0
run small_nofold small OXB_FOLD_HB=0
