#!/bin/bash
# round 2, call AF: oxDNA3 -- remaining GPU tests, drop-in executable against the reference CPU binary, bench of the C2 geometry under oxDNA3
# against the reference's own CUDA backend (interaction_type = DNA3), ncu of the new kernel
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_dna3.py tests/test_dropin.py -q -x -k "dna3" 2>&1 | tail -30 ) > gpurun_out/r2af_tests.log 2>&1
tail -3 gpurun_out/r2af_tests.log
timeout 900 python bench.py --workload c2_dna3 --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2af_c2_dna3.json 2> gpurun_out/r2af_c2_dna3.err
python - <<PY
import json
try:
    b=json.loads(open("gpurun_out/r2af_c2_dna3.json").read().strip().splitlines()[-1]); k=b.get("kernels_ms")
    print("c2_dna3", "%.4g" % b["value"], {x: round(k[x],4) for x in ("force_pass","integrate","list_build_per_rebuild","sort_per_sort","md_step_mean")}, "ref_cuda", (b.get("reference_cuda") or {}).get("value"), b["config"].get("pairs_per_particle"))
except Exception as e: print("c2_dna3 failed", e); print(open("gpurun_out/r2af_c2_dna3.err").read()[-1500:])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_forces_dna3 -s 200 -c 2 -o gpurun_out/prof_dna3_r02af -f python bench.py --workload c2_dna3 --steps 1 --warmup 1 --no-cpu-baseline --no-extras --no-ref-cuda > gpurun_out/ncu_dna3_r02af.log 2>&1
tail -2 gpurun_out/ncu_dna3_r02af.log
