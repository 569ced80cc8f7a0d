#!/bin/bash
# round 2, call AW: register caps of the oxDNA3 edge kernels (OXB_DNA3_EDGE_MB = 0 / 1 / 2) at 1M and 82k nucleotides
mkdir -p gpurun_out
( OXB_DNA3_EDGE_MB=2 timeout 600 python -m pytest tests/test_gpu_dna3.py -q -x -k "forces_torques or nicked or full_size" 2>&1 | tail -2 ) > gpurun_out/r2aw_tests.log 2>&1
tail -1 gpurun_out/r2aw_tests.log
Q="--no-cpu-baseline --no-extras --no-ref-cuda"
for cfg in 0 1 2; do
 for wl in c4_dna3 c2_dna3; do
  OXB_DNA3_EDGE_MB=$cfg timeout 600 python bench.py --workload $wl --steps 3 --warmup 2 --equil 3000 $Q > gpurun_out/r2aw_${wl}_$cfg.json 2> gpurun_out/r2aw_${wl}_$cfg.err
  python - <<PY
import json
try:
    b=json.loads(open("gpurun_out/r2aw_${wl}_$cfg.json").read().strip().splitlines()[-1]); k=b.get("kernels_ms")
    print("r2aw_${wl}_mb$cfg", "%.4g" % b["value"], {x: round(k[x],4) for x in ("force_pass","integrate","list_build_per_rebuild","sort_per_sort","md_step_mean")})
except Exception as e: print("r2aw_${wl}_$cfg failed", e)
PY
 done
done
