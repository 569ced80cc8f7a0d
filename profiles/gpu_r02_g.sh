#!/bin/bash
# round 2, call G: half-shell neighbour scan, plugin seam (C ABI callback + PluginManager fallback), full GPU suite
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 ) > gpurun_out/r2g_tests.log 2>&1
tail -5 gpurun_out/r2g_tests.log
Q="--no-cpu-baseline --no-ref-cuda --no-extras"
timeout 600 python bench.py --workload c4 --steps 3 --warmup 3 $Q > gpurun_out/r2g_c4.json 2> gpurun_out/r2g_c4.err
OXB_HALF_SHELL=0 timeout 600 python bench.py --workload c4 --steps 3 --warmup 3 $Q > gpurun_out/r2g_c4_full.json 2> gpurun_out/r2g_c4_full.err
timeout 600 python bench.py --workload c2 --steps 5 --warmup 3 $Q > gpurun_out/r2g_c2.json 2> gpurun_out/r2g_c2.err
OXB_BUILD_G=1 timeout 600 python bench.py --workload c2 --steps 5 --warmup 3 $Q > gpurun_out/r2g_c2_g1.json 2> gpurun_out/r2g_c2_g1.err
for f in r2g_c4 r2g_c4_full r2g_c2 r2g_c2_g1; do python - <<PY
import json
try:
    b=json.load(open("gpurun_out/$f.json")); k=b["kernels_ms"]; print("$f", "%.4g" % b["value"], {x: round(k[x],4) for x in ("force_pass","integrate","list_build_per_rebuild","sort_per_sort","md_step_mean")})
except Exception as e: print("$f", "failed", e)
PY
done
