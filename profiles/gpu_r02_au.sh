#!/bin/bash
# round 2, call AU: oxDNA3 at 1M nucleotides (C4 geometry, average-sequence tables) against the reference's CUDA backend with interaction_type = DNA3
mkdir -p gpurun_out
timeout 1200 python bench.py --workload c4_dna3 --steps 3 --warmup 3 --equil 3000 --no-cpu-baseline --no-extras > gpurun_out/r2au_c4_dna3.json 2> gpurun_out/r2au_c4_dna3.err
python - <<PY
import json
try:
    b=json.loads(open("gpurun_out/r2au_c4_dna3.json").read().strip().splitlines()[-1]); k=b.get("kernels_ms")
    rc=b.get("reference_cuda") or {}
    print("c4_dna3", "%.4g" % b["value"], "e2e", "%.4g" % b["e2e"]["value"], {x: round(k[x],4) for x in ("force_pass","integrate","list_build_per_rebuild","sort_per_sort","md_step_mean")}, "ref_cuda", rc.get("value"), [(r.get("use_edge"), r.get("CUDA_sort_every"), r.get("value")) for r in rc.get("runs", [])], (rc.get("no_timer_sync") or {}).get("value"))
except Exception as e: print("c4_dna3 failed", e); print(open("gpurun_out/r2au_c4_dna3.err").read()[-1500:])
PY
