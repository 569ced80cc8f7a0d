#!/bin/bash
# round 2, final pass: full GPU suite, the driver's default bench commands (ours + reference arm), launch list of the same command under ncu
# (cold-cache, serialised: compare SHARES), ncu --set full of every hot kernel at C4 and C2
mkdir -p gpurun_out
TAG=${1:-r02}
( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 ) > gpurun_out/tests_${TAG}.log 2>&1
tail -4 gpurun_out/tests_${TAG}.log
( python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 ) | tee gpurun_out/smoke_${TAG}.log
( time timeout 1500 python bench.py > gpurun_out/bench_${TAG}_default.json 2> gpurun_out/bench_${TAG}_default.err ) 2>&1 | tail -3
( time timeout 900 python bench.py --impl reference > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err ) 2>&1 | tail -3
Q="--no-cpu-baseline --no-ref-cuda --no-extras"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 40000 -c 900 --csv --log-file gpurun_out/launches_c4_${TAG}.csv \
  python bench.py --workload c4 --steps 1 --warmup 1 --equil 6000 --md-steps 100 $Q > gpurun_out/ncu_launches_c4_${TAG}.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 30000 -c 900 --csv --log-file gpurun_out/launches_c2_${TAG}.csv \
  python bench.py --workload c2 --steps 1 --warmup 1 --equil 6000 --md-steps 100 $Q > gpurun_out/ncu_launches_c2_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_edge|k_bonded|k_dh|k_integrate|k_build_neigh|k_fill_edges|k_permute|k_sort_small|k_ext" -s 700 -c 14 -o gpurun_out/prof_c4_${TAG} -f \
    python bench.py --workload c4 --steps 1 --warmup 1 --md-steps 30 --equil 400 $Q > gpurun_out/ncu_full_c4_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_edge|k_bonded|k_dh|k_integrate|k_build_neigh|k_fill_edges|k_permute|k_sort_small" -s 600 -c 14 -o gpurun_out/prof_c2_${TAG} -f \
    python bench.py --workload c2 --steps 1 --warmup 1 --md-steps 60 --equil 400 $Q > gpurun_out/ncu_full_c2_${TAG}.log 2>&1
ls -la gpurun_out/*${TAG}* | tail -14
