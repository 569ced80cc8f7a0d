#!/bin/bash
# round 2, closing pass on the final code: full GPU suite, smoke, the driver's default bench commands (ours + reference arm)
mkdir -p gpurun_out
TAG=${1:-r02e}
( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 ) > gpurun_out/tests_${TAG}.log 2>&1
tail -5 gpurun_out/tests_${TAG}.log
( python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 ) | tee gpurun_out/smoke_${TAG}.log
( time timeout 1800 python bench.py > gpurun_out/bench_${TAG}_default.json 2> gpurun_out/bench_${TAG}_default.err ) 2>&1 | tail -3
( time timeout 900 python bench.py --impl reference > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err ) 2>&1 | tail -3
python - <<PY
import json
b=json.loads(open("gpurun_out/bench_${TAG}_default.json").read().strip().splitlines()[-1])
print("C4", "%.4g" % b["value"], "e2e", "%.4g" % b["e2e"]["value"], "ref_cuda", (b.get("reference_cuda") or {}).get("value"))
for k, v in b.get("extras", {}).items(): print(k, v.get("value"), (v.get("reference_cuda") or {}).get("value") if isinstance(v.get("reference_cuda"), dict) else None, v.get("error"))
PY
