#!/bin/bash
# round 2, call AT: oxDNA3 near-edge kernel screens the work lists on the per-record radial gates
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_dna3.py -q -x 2>&1 | tail -25 ) > gpurun_out/r2at_tests.log 2>&1
tail -3 gpurun_out/r2at_tests.log
Q="--no-cpu-baseline --no-extras --no-ref-cuda"
for ue in 1; do
  timeout 600 python bench.py --workload c2_dna3 --use-edge $ue --steps 3 --warmup 3 $Q > gpurun_out/r2at_ue$ue.json 2> gpurun_out/r2at_ue$ue.err
  python - <<PY
import json
try:
    b=json.loads(open("gpurun_out/r2at_ue$ue.json").read().strip().splitlines()[-1]); k=b.get("kernels_ms")
    print("r2at_ue$ue", "%.4g" % b["value"], {x: round(k[x],4) for x in ("force_pass","integrate","list_build_per_rebuild","sort_per_sort","md_step_mean")}, b["config"]["pairs_per_particle"])
except Exception as e: print("r2at_ue$ue failed", e); print(open("gpurun_out/r2at_ue$ue.err").read()[-600:])
PY
done
