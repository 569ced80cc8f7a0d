#!/bin/bash
# round 2, call AM (8 GPUs): C5 on eight and on four GPUs (the driver's multi-GPU command)
mkdir -p gpurun_out
for n in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/r2am_c5_n$n.json 2> gpurun_out/r2am_c5_n$n.err
python - <<PY
import json
try:
    b=json.loads([l for l in open("gpurun_out/r2am_c5_n$n.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("c5_n$n", "%.4g" % b["value"], "e2e", b.get("e2e",{}).get("value"), [(round(r.get("md_ms",0)), round(r.get("exchange_ms",0),1), round(r.get("host_wait_ms",0),1)) for r in b.get("per_rank",[])])
except Exception as e: print("c5_n$n failed", e); print(open("gpurun_out/r2am_c5_n$n.err").read()[-800:])
PY
done
