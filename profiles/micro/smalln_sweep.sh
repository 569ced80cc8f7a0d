#!/bin/bash
# small-system tuning (C2, 81,920 nt): lanes per particle in the Debye-Hueckel kernel, block sizes of integrate / bonded; C4 as a guard
mkdir -p gpurun_out
run() { # label, workload args...
  local label="$1"; shift
  python bench.py "$@" --no-ref-cuda --no-cpu-baseline > gpurun_out/sw.json 2> gpurun_out/sw.err
  python -c "
import json; d=json.load(open('gpurun_out/sw.json')); k=d['kernels_ms']; print('$label  value %.3e forces_ms %.4f integrate_ms %.4f step %.4f' % (d['value'], k['forces'], k['integrate'], k['md_step_mean']))"
}
{
for lpp in 1 2 4; do OXB_DH_LPP=$lpp run "c2 dh_lpp=$lpp" --md-steps 1000 --steps 4 --warmup 3; done
for t in 256 128 64; do OXB_TPB_INTEGRATE=$t run "c2 tpb_integrate=$t" --md-steps 1000 --steps 4 --warmup 3; done
for t in 128 64; do OXB_TPB_BONDED=$t run "c2 tpb_bonded=$t" --md-steps 1000 --steps 4 --warmup 3; done
OXB_DH_LPP=4 OXB_TPB_INTEGRATE=128 OXB_TPB_BONDED=64 run "c2 lpp=4 integ=128 bonded=64" --md-steps 1000 --steps 4 --warmup 3
for lpp in 1 2 4; do OXB_DH_LPP=$lpp run "c4 dh_lpp=$lpp" --workload c4 --md-steps 200 --steps 3 --warmup 3 --equil 600; done
OXB_TPB_INTEGRATE=128 run "c4 tpb_integrate=128" --workload c4 --md-steps 200 --steps 3 --warmup 3 --equil 600
} 2>&1 | tee gpurun_out/smalln_sweep.log
