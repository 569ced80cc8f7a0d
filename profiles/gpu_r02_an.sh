#!/bin/bash
# round 2, call AN: oxDNA3 with records fetched where they are used (fewer live registers); one launch against two (OXB_DNA3_TWO) at 168 / 128 registers
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_dna3.py -q -x 2>&1 | tail -5 ) > gpurun_out/r2an_tests.log 2>&1
tail -1 gpurun_out/r2an_tests.log
( OXB_DNA3_TWO=1 OXB_DNA3_MB=4 timeout 900 python -m pytest tests/test_gpu_dna3.py -q -x 2>&1 | tail -5 ) > gpurun_out/r2an_tests_two.log 2>&1
tail -1 gpurun_out/r2an_tests_two.log
Q="--no-cpu-baseline --no-extras --no-ref-cuda"
for cfg in "0 3" "0 4" "1 4" "1 3"; do
  set -- $cfg
  OXB_DNA3_TWO=$1 OXB_DNA3_MB=$2 timeout 600 python bench.py --workload c2_dna3 --steps 3 --warmup 3 $Q > gpurun_out/r2an_two$1_mb$2.json 2> gpurun_out/r2an_two$1_mb$2.err
  python - <<PY
import json
try:
    b=json.loads(open("gpurun_out/r2an_two$1_mb$2.json").read().strip().splitlines()[-1]); k=b.get("kernels_ms")
    print("r2an_two$1_mb$2", "%.4g" % b["value"], {x: round(k[x],4) for x in ("force_pass","integrate","list_build_per_rebuild","sort_per_sort","md_step_mean")})
except Exception as e: print("r2an_two$1_mb$2 failed", e)
PY
done
