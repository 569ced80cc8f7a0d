# round 1: launch-bounds sweep for the integrator (C4, edge mode) and the particle-centric force kernel (C2, use_edge = 0)
mkdir -p gpurun_out
for v in 1 5 6 8; do
  OXB_EXTRA_NVCC="-DOXB_MB_INTEGRATE=$v" python oxdna_b200/build.py --force > /dev/null 2>&1
  python bench.py --workload c4 --md-steps 100 --steps 3 --warmup 3 --equil 400 --no-cpu-baseline --no-ref-cuda > gpurun_out/occ.json 2> gpurun_out/occ.err
  python -c "
import json; d=json.load(open('gpurun_out/occ.json')); print('integrate min blocks = $v  c4 value %.4g integrate_ms %.4f step %.4f' % (d['value'], d['kernels_ms']['integrate'], d['kernels_ms']['md_step_mean']))" | tee -a gpurun_out/occupancy_sweep2.log
done
for v in 1 4 5 6; do
  OXB_EXTRA_NVCC="-DOXB_MB_PARTICLE=$v" python oxdna_b200/build.py --force > /dev/null 2>&1
  python bench.py --workload c2 --use-edge 0 --md-steps 1000 --steps 3 --warmup 3 --equil 5000 --no-cpu-baseline --no-ref-cuda > gpurun_out/occ.json 2> gpurun_out/occ.err
  python -c "
import json; d=json.load(open('gpurun_out/occ.json')); print('particle-centric min blocks = $v  c2 value %.4g forces_ms %.4f step %.4f' % (d['value'], d['kernels_ms']['forces'], d['kernels_ms']['md_step_mean']))" | tee -a gpurun_out/occupancy_sweep2.log
done
