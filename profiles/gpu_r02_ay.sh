#!/bin/bash
# round 2, call AY: full GPU suite after the dummy-base handling (btype = type = 4) went into the shared type function; short C4 / C2 bench for regressions
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 ) > gpurun_out/r2ay_tests.log 2>&1
tail -5 gpurun_out/r2ay_tests.log
Q="--no-cpu-baseline --no-ref-cuda --no-extras"
for wl in c4 c2; do
  timeout 600 python bench.py --workload $wl --steps 3 --warmup 3 $Q > gpurun_out/r2ay_$wl.json 2> gpurun_out/r2ay_$wl.err
  python - <<PY
import json
try:
    b=json.loads(open("gpurun_out/r2ay_$wl.json").read().strip().splitlines()[-1]); k=b.get("kernels_ms")
    print("r2ay_$wl", "%.4g" % b["value"], {x: round(k[x],4) for x in ("force_pass","integrate","list_build_per_rebuild","sort_per_sort","md_step_mean")})
except Exception as e: print("r2ay_$wl failed", e)
PY
done
