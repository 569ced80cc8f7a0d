#!/bin/bash
# round 2, call AG: oxDNA3 kernel in cost-split passes; register cap sweep (OXB_DNA3_MB = 2, 3, 4); ncu of the default
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_dna3.py -q -x 2>&1 | tail -5 ) > gpurun_out/r2ag_tests.log 2>&1
tail -1 gpurun_out/r2ag_tests.log
Q="--no-cpu-baseline --no-extras --no-ref-cuda"
for mb in 3 2 4; do
  OXB_DNA3_MB=$mb timeout 600 python bench.py --workload c2_dna3 --steps 3 --warmup 3 $Q > gpurun_out/r2ag_mb$mb.json 2> gpurun_out/r2ag_mb$mb.err
  python - <<PY
import json
try:
    b=json.loads(open("gpurun_out/r2ag_mb$mb.json").read().strip().splitlines()[-1]); k=b.get("kernels_ms")
    print("r2ag_mb$mb", "%.4g" % b["value"], {x: round(k[x],4) for x in ("force_pass","integrate","list_build_per_rebuild","sort_per_sort","md_step_mean")})
except Exception as e: print("r2ag_mb$mb failed", e)
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_forces_dna3 -s 200 -c 1 -o gpurun_out/prof_dna3_r02ag -f python bench.py --workload c2_dna3 --steps 1 --warmup 1 $Q > gpurun_out/ncu_dna3_r02ag.log 2>&1
tail -1 gpurun_out/ncu_dna3_r02ag.log | cut -c1-80
