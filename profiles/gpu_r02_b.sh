#!/bin/bash
# round 2, call B: suite after the half-matrix DH / packed staleness references; C4 A/B; ncu launch list at C4; C5 per-GPU load
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 ) > gpurun_out/r2b_tests.log 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "rna_long_run" 2>&1 | tail -8 > gpurun_out/r2b_rna_stat2.log
Q="--no-cpu-baseline --no-ref-cuda --no-extras"
timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 $Q > gpurun_out/r2b_c4_half.json 2> gpurun_out/r2b_c4_half.err
OXB_DH_HALF=0 timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 $Q > gpurun_out/r2b_c4_full.json 2> gpurun_out/r2b_c4_full.err
OXB_FOLD=0 timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 $Q > gpurun_out/r2b_c4_nofold.json 2> gpurun_out/r2b_c4_nofold.err
timeout 600 python bench.py --workload c2 --steps 5 --warmup 3 $Q > gpurun_out/r2b_c2_half.json 2> gpurun_out/r2b_c2_half.err
OXB_FOLD=0 timeout 600 python bench.py --workload c2 --steps 5 --warmup 3 $Q > gpurun_out/r2b_c2_nofold.json 2> gpurun_out/r2b_c2_nofold.err
timeout 600 python bench.py --workload c5 --replicas 8 --steps 3 --warmup 2 --equil 5000 $Q > gpurun_out/r2b_c5_8rep.json 2> gpurun_out/r2b_c5_8rep.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 600 --csv --log-file gpurun_out/r2b_launches_c4.csv \
  python bench.py --workload c4 --steps 1 --warmup 1 --equil 400 --md-steps 100 $Q > gpurun_out/r2b_ncu_c4.log 2>&1
tail -3 gpurun_out/r2b_tests.log; tail -3 gpurun_out/r2b_rna_stat2.log
for f in r2b_c4_half r2b_c4_full r2b_c4_nofold r2b_c2_half r2b_c2_nofold r2b_c5_8rep; do python - <<PY
import json
try:
    b=json.load(open("gpurun_out/$f.json")); print("$f", "%.4g" % b["value"], json.dumps(b.get("kernels_ms"))[:400])
except Exception as e: print("$f", "failed", e)
PY
done
