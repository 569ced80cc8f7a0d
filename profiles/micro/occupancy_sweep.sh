# round 1: sweep of __launch_bounds__ min-blocks for the gather kernels; rebuilds the library on the GPU box per variant and times the C4 force pass
mkdir -p gpurun_out
for v in "1 1 1" "6 1 1" "8 1 1" "1 12 1" "1 16 1" "1 1 6" "1 1 8" "6 12 6" "8 16 8"; do
  set -- $v
  OXB_EXTRA_NVCC="-DOXB_MB_NEAR=$1 -DOXB_MB_HEAVY=$2 -DOXB_MB_BONDED=$3" python oxdna_b200/build.py --force > /dev/null 2>&1
  python bench.py --workload c4 --md-steps 100 --steps 2 --warmup 3 --equil 400 --no-cpu-baseline --no-ref-cuda > gpurun_out/occ.json 2> gpurun_out/occ.err
  python -c "
import json; d=json.load(open('gpurun_out/occ.json')); print('near/heavy/bonded min blocks = $v  C4 value %.4g forces_ms %.4f step %.4f' % (d['value'], d['kernels_ms']['forces'], d['kernels_ms']['md_step_mean']))" | tee -a gpurun_out/occupancy_sweep.log
done
