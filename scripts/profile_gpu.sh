#!/bin/bash
# Run under gpurun on ONE GPU.  Produces the ncu launch list of the bench command and one full capture of the
# two dominant kernels; outputs land in gpurun_out/ (scratch) and are summarised into profiles/ by
# scripts/summarise_profiles.py on the CPU box.
set -x
mkdir -p gpurun_out
TAG=${1:-r01}
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_(optimize|spmv|pack|build|stiefel)" -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 12 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_spmv -s 3 -c 2 -f -o gpurun_out/prof_spmv_${TAG} \
    python bench.py --steps 6 --warmup 3 --no-cpu > gpurun_out/ncu_spmv_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_optimize -s 10 -c 1 -f -o gpurun_out/prof_opt_${TAG} \
    python bench.py --steps 6 --warmup 3 --no-cpu --no-spmv > gpurun_out/ncu_opt_${TAG}.log 2>&1
ls -la gpurun_out
