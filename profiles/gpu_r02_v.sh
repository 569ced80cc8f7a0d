#!/bin/bash
# round 2, call V: near-edge classes (the list builder tells the near-edge kernel which families of site pairs can come into range); full GPU suite
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 ) > gpurun_out/r2v_tests.log 2>&1
tail -4 gpurun_out/r2v_tests.log
( OXB_BUILD_G=1 timeout 900 python -m pytest tests -m gpu -q -x -k "forces_torques or rna_forces or full_size_c2 or replica" 2>&1 | tail -2 ) > gpurun_out/r2v_tests_g1.log 2>&1
tail -1 gpurun_out/r2v_tests_g1.log
Q="--no-cpu-baseline --no-ref-cuda --no-extras"
run() { tag=$1; wl=$2; shift 2
  env "$@" timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 $Q > gpurun_out/r2v_$tag.json 2> gpurun_out/r2v_$tag.err
  python - <<PY
import json
try:
    b=json.load(open("gpurun_out/r2v_$tag.json")); k=b["kernels_ms"]; print("r2v_$tag", "%.4g" % b["value"], {x: round(k[x],4) for x in ("force_pass","integrate","list_build_per_rebuild","sort_per_sort","md_step_mean")})
except Exception as e: print("r2v_$tag", "failed", e)
PY
}
run c4 c4 X=0
run c2 c2 X=0
run c3 c3 X=0
