#!/bin/bash
# round 2, call X (8 GPUs): the driver's N = 8 and N = 4 commands (C5: 64 replicas, 8 / 16 per GPU)
mkdir -p gpurun_out
for N in ${NS:-8 4}; do
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N > gpurun_out/r2x_c5_n$N.json 2> gpurun_out/r2x_c5_n$N.err ) 2>&1 | tail -3
python - <<PY
import json
try:
    b=json.loads(open("gpurun_out/r2x_c5_n$N.json").read().strip().splitlines()[-1]); print("n$N", "%.4g" % b["value"], b["ms_per_step"], [ (r["rank"], round(r["md_ms"],1), round(r["exchange_ms"],2), round(r["host_wait_ms"],2)) for r in b["per_rank"]], "e2e %.4g" % b["e2e"]["value"], b["config"]["exchange_acceptance"])
except Exception as e: print("n$N failed", e)
PY
done
