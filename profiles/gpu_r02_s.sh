#!/bin/bash
# round 2, call S: replica batch of 8 (the per-GPU load of C5 on 8 GPUs): which of the round's changes pay there
mkdir -p gpurun_out
Q="--no-cpu-baseline --no-ref-cuda --no-extras --replicas 8"
run() { tag=$1; shift
  env "$@" timeout 600 python bench.py --workload c5 --steps 3 --warmup 2 $Q > gpurun_out/r2s_$tag.json 2> gpurun_out/r2s_$tag.err
  python - <<PY
import json
try:
    b=json.loads(open("gpurun_out/r2s_$tag.json").read().strip().splitlines()[-1]); print("r2s_$tag", "%.4g" % b["value"], b["per_rank"][0]["list_rebuilds_per_replica_batch"])
except Exception as e: print("r2s_$tag", "failed", e)
PY
}
run base X=0
run full OXB_HALF_SHELL=0
run g8 OXB_BUILD_G=8
run nofork OXB_FORK=0
run foldhb OXB_FOLD_HB=1
