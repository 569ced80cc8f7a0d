#!/bin/bash
# round 2, call M: ncu --set full with source of the force-pass kernels and the integrator at C4 (thermalised), + short C4 bench of the current code
mkdir -p gpurun_out
Q="--no-cpu-baseline --no-ref-cuda --no-extras"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_edge|k_bonded|k_dh|k_integrate" -s 600 -c 5 -o gpurun_out/prof_force_r02m -f \
    python bench.py --workload c4 --steps 1 --warmup 1 --md-steps 30 --equil 400 $Q > gpurun_out/ncu_force_r02m.log 2>&1
timeout 600 python bench.py --workload c4 --steps 3 --warmup 3 $Q > gpurun_out/r2m_c4.json 2> gpurun_out/r2m_c4.err
python - <<PY
import json
b=json.load(open("gpurun_out/r2m_c4.json")); k=b["kernels_ms"]; print("r2m_c4", "%.4g" % b["value"], {x: round(k[x],4) for x in ("force_pass","integrate","list_build_per_rebuild","sort_per_sort","md_step_mean")}, {x: round(v,4) for x,v in k["rebuild_parts"].items()})
PY
ls -la gpurun_out/*r02m*
