# round 1: sweep of __launch_bounds__ min-blocks for the gather kernels; rebuilds the library on the GPU box per variant and times the force pass
# usage: occupancy_sweep.sh "<near heavy bonded>;<...>" "<workloads>"
mkdir -p gpurun_out
WORKLOADS="${2:-c4}"
IFS=";" read -ra VARS <<< "${1:-1 1 1;8 16 8}"
for v in "${VARS[@]}"; do
  set -- $v
  OXB_EXTRA_NVCC="-DOXB_MB_NEAR=$1 -DOXB_MB_HEAVY=$2 -DOXB_MB_BONDED=$3" python oxdna_b200/build.py --force > /dev/null 2>&1
  for w in $WORKLOADS; do
    if [ "$w" = "c4" ]; then A="--md-steps 100 --steps 3 --warmup 3 --equil 400"; else A="--md-steps 1000 --steps 3 --warmup 3 --equil 5000"; fi
    python bench.py --workload $w $A --no-cpu-baseline --no-ref-cuda > gpurun_out/occ.json 2> gpurun_out/occ.err
    python -c "
import json; d=json.load(open('gpurun_out/occ.json')); print('near/heavy/bonded min blocks = $v  $w value %.4g forces_ms %.4f step %.4f' % (d['value'], d['kernels_ms']['forces'], d['kernels_ms']['md_step_mean']))" | tee -a gpurun_out/occupancy_sweep.log
  done
done
