#!/bin/bash
# round 2, call C: launch-gap phase, thermalised ncu launch list at C4, fold / nofold, C5 per-GPU load
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 ) > gpurun_out/r2c_tests.log 2>&1
Q="--no-cpu-baseline --no-ref-cuda --no-extras"
timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 $Q > gpurun_out/r2c_c4.json 2> gpurun_out/r2c_c4.err
OXB_NO_GRAPHS=1 timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 $Q > gpurun_out/r2c_c4_nograph.json 2> gpurun_out/r2c_c4_nograph.err
OXB_FORK=0 timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 $Q > gpurun_out/r2c_c4_nofork.json 2> gpurun_out/r2c_c4_nofork.err
timeout 600 python bench.py --workload c2 --steps 5 --warmup 3 $Q > gpurun_out/r2c_c2.json 2> gpurun_out/r2c_c2.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 40000 -c 700 --csv --log-file gpurun_out/r2c_launches_c4.csv \
  python bench.py --workload c4 --steps 1 --warmup 1 --equil 6000 --md-steps 100 $Q > gpurun_out/r2c_ncu_c4.log 2>&1
tail -3 gpurun_out/r2c_tests.log
for f in r2c_c4 r2c_c4_nograph r2c_c4_nofork r2c_c2; do python - <<PY
import json
try:
    b=json.load(open("gpurun_out/$f.json")); print("$f", "%.4g" % b["value"], json.dumps(b.get("kernels_ms"))[:600])
except Exception as e: print("$f", "failed", e)
PY
done
