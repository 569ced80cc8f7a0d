#!/bin/bash
# Launch list (device time per kernel) of the UNMODIFIED reference CUDA backend on C2 / C4, for comparison of where time goes.
set -x
mkdir -p gpurun_out /tmp/refprof
python - <<PY
import sys, os
sys.path.insert(0, os.getcwd())
import bench
from oracle import ref_cuda_bench as R
from oxdna_b200 import lattice
for w, steps, ext in (("c2", 600, False), ("c4", 200, True)):
    sysm, desc = bench.workload(w)
    d = f"/tmp/refprof/{w}"
    os.makedirs(d, exist_ok=True)
    top, conf = bench.write_case(sysm, bench.T_STR, d)
    ext_path = None
    if ext:
        ext_path = os.path.join(d, "forces.txt")
        R.write_forces_file(ext_path, lattice.mutual_traps(sysm))
    open(os.path.join(d, "input"), "w").write(R.TEMPLATE.format(salt=0.5, T="300K", dt=0.003, steps=steps, sort_every=0 if ext else 1, use_edge=1, top=top, conf=conf, d=d,
        ext=1 if ext else 0, extfile=f"external_forces_file = {ext_path}" if ext_path else ""))
PY
for w in c2 c4; do
  ( cd /tmp/refprof/$w && ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 500 --csv --log-file $OLDPWD/gpurun_out/launches_ref_$w.csv $OLDPWD/oracle/_ref/oxDNA_cuda input > $OLDPWD/gpurun_out/ncu_ref_$w.log 2>&1 )
done
ls -la gpurun_out
