#!/bin/bash
# Run under gpurun on ONE GPU.  Produces the ncu launch list of the bench command and one full capture of the dominant
# kernels; outputs land in gpurun_out/ (scratch) and are summarised into profiles/ by scripts/summarise_profiles.py on
# the CPU box.   usage: scripts/profile_gpu.sh r02
set -x
mkdir -p gpurun_out
TAG=${1:-r02}
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_(optimize|spmv|pack|build|stiefel)" -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 12 --warmup 3 --no-cpu --no-sweep --no-multi > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_spmv -s 3 -c 1 -f -o gpurun_out/prof_spmv_${TAG} \
    python bench.py --steps 6 --warmup 3 --no-cpu --no-sweep --no-multi > gpurun_out/ncu_spmv_${TAG}.log 2>&1
# the persistent step kernel with the default (nested-dissection) exact preconditioner, and the dense inverse for A/B
ncu --set full --clock-control none --import-source on -k regex:k_optimize -s 10 -c 1 -f -o gpurun_out/prof_opt_${TAG} \
    python bench.py --steps 6 --warmup 3 --no-cpu --no-spmv --no-multi > gpurun_out/ncu_opt_${TAG}.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_optimize -s 10 -c 1 --csv \
    --log-file gpurun_out/traffic_opt_dense_${TAG}.csv python bench.py --steps 6 --warmup 3 --no-cpu --no-spmv --no-multi --precond dense \
    > gpurun_out/ncu_opt_dense_${TAG}.log 2>&1
ls -la gpurun_out
