#!/bin/bash
# round 2, call Z: tile-staged variant of the near-edge kernel (OXB_NEAR_TILE=1: 128 `from` slots per block staged in shared memory)
mkdir -p gpurun_out
( OXB_NEAR_TILE=1 timeout 900 python -m pytest tests -m gpu -q -x -k "forces_torques or rna_forces or full_size_c2 or nve or work_list" 2>&1 | tail -3 ) > gpurun_out/r2z_tests.log 2>&1
tail -1 gpurun_out/r2z_tests.log
Q="--no-cpu-baseline --no-ref-cuda --no-extras"
run() { tag=$1; wl=$2; shift 2
  env "$@" timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 $Q > gpurun_out/r2z_$tag.json 2> gpurun_out/r2z_$tag.err
  python - <<PY
import json
try:
    b=json.load(open("gpurun_out/r2z_$tag.json")); k=b["kernels_ms"]; print("r2z_$tag", "%.4g" % b["value"], {x: round(k[x],4) for x in ("force_pass","integrate","list_build_per_rebuild","sort_per_sort","md_step_mean")})
except Exception as e: print("r2z_$tag", "failed", e)
PY
}
run c4 c4 X=0
run c4_tile c4 OXB_NEAR_TILE=1
run c2 c2 X=0
run c2_tile c2 OXB_NEAR_TILE=1
run c3_tile c3 OXB_NEAR_TILE=1
timeout 600 ncu --set full --clock-control none -k regex:"k_edge_near" -s 300 -c 1 -o gpurun_out/prof_near_tile_r02z -f \
    env OXB_NEAR_TILE=1 python bench.py --workload c4 --steps 1 --warmup 1 --md-steps 30 --equil 400 $Q > gpurun_out/ncu_near_tile_r02z.log 2>&1
