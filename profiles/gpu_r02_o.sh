#!/bin/bash
# round 2, call O (2 GPUs): the driver's N = 2 command (C5: 64 replicas, 32 per GPU) and the 2-rank reference arm
mkdir -p gpurun_out
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r2o_c5_n2.json 2> gpurun_out/r2o_c5_n2.err ) 2>&1 | tail -3
python - <<PY
import json
try:
    b=json.loads(open("gpurun_out/r2o_c5_n2.json").read().strip().splitlines()[-1]); print("n2", "%.4g" % b["value"], b["ms_per_step"], b["per_rank"], b["e2e"], b["config"]["exchange_acceptance"], b["config"].get("batches_per_gpu"))
except Exception as e: print("failed", e)
PY
tail -5 gpurun_out/r2o_c5_n2.err
