#!/bin/bash
# round 2, call AL (2 GPUs): the full-size oxDNA3 test; C5 on two GPUs (the driver's multi-GPU command) after the round's changes
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_dna3.py -q -x -k full_size 2>&1 | tail -5 ) > gpurun_out/r2al_tests.log 2>&1
tail -1 gpurun_out/r2al_tests.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2al_c5_n2.json 2> gpurun_out/r2al_c5_n2.err
python - <<PY
import json
try:
    b=json.loads([l for l in open("gpurun_out/r2al_c5_n2.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("c5_n2", "%.4g" % b["value"], "e2e", b.get("e2e",{}).get("value"), b["config"].get("workload","")[:60])
except Exception as e: print("c5_n2 failed", e); print(open("gpurun_out/r2al_c5_n2.err").read()[-800:])
PY
