#!/bin/bash
# round 2, call AQ: forked streams for the force pass at 1M nt re-measured with the round's kernels (OXB_FORK = 0 / 1); initcheck after zeroing the matrix
mkdir -p gpurun_out
( timeout 600 compute-sanitizer --tool initcheck --print-limit 10 python -m pytest tests/test_gpu_dna3.py -q -x -k "forces_torques and lattice8 and mixed and 0-0" 2>&1 | grep -E "Uninitialized|ERROR SUMMARY|passed|failed" | head -8 ) > gpurun_out/r2aq_init_dna3.log 2>&1
tail -2 gpurun_out/r2aq_init_dna3.log
Q="--no-cpu-baseline --no-ref-cuda --no-extras"
run() { tag=$1; wl=$2; shift 2
  env "$@" timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 $Q > gpurun_out/r2aq_$tag.json 2> gpurun_out/r2aq_$tag.err
  python - <<PY
import json
try:
    b=json.loads(open("gpurun_out/r2aq_$tag.json").read().strip().splitlines()[-1]); k=b.get("kernels_ms")
    print("r2aq_$tag", "%.4g" % b["value"], {x: round(k[x],4) for x in ("force_pass","integrate","list_build_per_rebuild","sort_per_sort","md_step_mean")} if k else "")
except Exception as e: print("r2aq_$tag", "failed", e)
PY
}
run c4_fork1 c4 OXB_FORK=1
run c4_fork0 c4 OXB_FORK=0
